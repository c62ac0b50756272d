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
#include <chrono>
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

#ifndef MPI_VERSION
// one process per GPU; without an MPI installation the communicator argument is a placeholder
using MPI_Comm = int;
constexpr MPI_Comm MPI_COMM_SELF = 1, MPI_COMM_WORLD = 0;
#endif

// The part of dealii::IndexSet the reference's SparseMatrixDevice constructor reads: a contiguous range of locally owned
// indices inside a global index space (sparse_matrix_device.templates.cuh:244-272 uses n_elements() and size()).
class IndexSet
{
public:
  using size_type = std::size_t;
  IndexSet() = default;
  explicit IndexSet(size_type size) : _size(size) {}
  void set_size(size_type size) { _size = size; }
  void add_range(size_type begin, size_type end)
  {
    ASSERT_THROW(begin <= end && end <= _size, "IndexSet::add_range: range outside the index space");
    ASSERT_THROW(_begin == _end || begin == _end, "IndexSet: only one contiguous range (one row block per rank)");
    if (_begin == _end)
      _begin = begin;
    _end = end;
  }
  void add_index(size_type index) { add_range(index, index + 1); }
  void compress() const {}
  size_type size() const { return _size; }
  size_type n_elements() const { return _end - _begin; }
  size_type nth_index_in_set(size_type n) const { return _begin + n; }
  bool is_element(size_type i) const { return i >= _begin && i < _end; }
  bool is_contiguous() const { return true; }

private:
  size_type _size = 0, _begin = 0, _end = 0;
};

// include/mfmg/cuda/sparse_matrix_device.cuh:28-104.  The reference's ctor takes an MPI_Comm and two
// dealii::IndexSet; without deal.II the (local) sizes are passed directly.
template <typename ScalarType>
class SparseMatrixDevice
{
  static_assert(sizeof(ScalarType) == sizeof(double), "FP64 only on this path");

public:
  // empty matrix, to be filled by reinit (sparse_matrix_device.cuh:30, the object a MeshEvaluator receives)
  SparseMatrixDevice() = default;
  // TAKES OWNERSHIP like the constructor below, releasing what the object held (sparse_matrix_device.templates.cuh:274-300)
  void reinit(std::shared_ptr<CudaHandle const> handle, ScalarType *val_dev_, int *column_index_dev_, int *row_ptr_dev_,
              unsigned int local_nnz, unsigned int n_rows, unsigned int n_cols)
  {
    release();
    _handle = std::move(handle);
    val_dev = val_dev_;
    column_index_dev = column_index_dev_;
    row_ptr_dev = row_ptr_dev_;
    check_status(_handle->ctx, mfmgb_csr_adopt_device(_handle->ctx, n_rows, n_cols, local_nnz, val_dev, column_index_dev,
                                                      row_ptr_dev, &_csr));
  }
  // convert_matrix semantics: copy a host CSR into this object
  void reinit(std::shared_ptr<CudaHandle const> handle, unsigned int n_rows, unsigned int n_cols,
              std::vector<int64_t> const &row_ptr, std::vector<int> const &column_index, std::vector<double> const &val)
  {
    release();
    _handle = std::move(handle);
    check_status(_handle->ctx, mfmgb_csr_upload(_handle->ctx, n_rows, n_cols, row_ptr.data(), column_index.data(),
                                                val.data(), &_csr));
    void *rp = nullptr;
    int is64 = 0;
    mfmgb_csr_device_arrays(_csr, &val_dev, &column_index_dev, &rp, &is64);
    row_ptr_dev = is64 ? nullptr : static_cast<int *>(rp);
  }
  // TAKES OWNERSHIP of the three cudaMalloc'ed arrays and frees them in the destructor
  // (sparse_matrix_device.templates.cuh:244-272)
  SparseMatrixDevice(std::shared_ptr<CudaHandle const> handle, ScalarType *val_dev_, int *column_index_dev_,
                     int *row_ptr_dev_, unsigned int local_nnz, unsigned int n_rows, unsigned int n_cols)
      : val_dev(val_dev_), column_index_dev(column_index_dev_), row_ptr_dev(row_ptr_dev_), _handle(std::move(handle))
  {
    check_status(_handle->ctx, mfmgb_csr_adopt_device(_handle->ctx, n_rows, n_cols, local_nnz, val_dev,
                                                      column_index_dev, row_ptr_dev, &_csr));
  }
  // the reference's own argument list (sparse_matrix_device.cuh:37-47): communicator + range / domain index sets, the
  // cuSPARSE handle replaced by the CudaHandle that owns the mfmgb context.  One rank owning every row is served here;
  // row-partitioned operators go through the halo-plan entry points of the C ABI (mfmgb_halo_create,
  // mfmgb_hierarchy_set_halo), where the column indices are LOCAL [owned | ghost].
  SparseMatrixDevice(MPI_Comm comm, ScalarType *val_dev_, int *column_index_dev_, int *row_ptr_dev_,
                     unsigned int local_nnz, IndexSet const &range_indexset, IndexSet const &domain_indexset,
                     std::shared_ptr<CudaHandle const> handle)
      : val_dev(val_dev_), column_index_dev(column_index_dev_), row_ptr_dev(row_ptr_dev_), _handle(std::move(handle)),
        _comm(comm)
  {
    ASSERT_THROW(range_indexset.n_elements() == range_indexset.size(),
                 "SparseMatrixDevice(MPI_Comm, ...): a rank that owns a row block only -- use the partitioned C ABI");
    check_status(_handle->ctx,
                 mfmgb_csr_adopt_device(_handle->ctx, (int64_t)range_indexset.n_elements(),
                                        (int64_t)domain_indexset.size(), local_nnz, val_dev, column_index_dev,
                                        row_ptr_dev, &_csr));
  }
  MPI_Comm get_mpi_communicator() const { return _comm; }
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
  ~SparseMatrixDevice() { release(); }
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

  // C = (*this) * B; C's previous content is released (sparse_matrix_device.templates.cuh:373-433; the reference uses
  // cusparseDcsrgemm on one rank and Trilinos on the host in parallel).  Setup operation, formed on the host here.
  void mmult(SparseMatrixDevice<ScalarType> &C, SparseMatrixDevice<ScalarType> const &B) const;

  mfmgb_csr *c_handle() const { return _csr; }
  std::shared_ptr<CudaHandle const> const &handle() const { return _handle; }

  // host copy of the CSR arrays (setup-time products, tests/test_utils_device.cu:221-263)
  void copy_to_host(std::vector<int64_t> &row_ptr, std::vector<int> &column_index, std::vector<double> &val) const
  {
    row_ptr.resize((std::size_t)info(0) + 1);
    column_index.resize((std::size_t)std::max<int64_t>(info(2), 1));
    val.resize((std::size_t)std::max<int64_t>(info(2), 1));
    check_status(_handle->ctx, mfmgb_csr_download(_handle->ctx, _csr, row_ptr.data(), column_index.data(), val.data()));
    column_index.resize((std::size_t)info(2));
    val.resize((std::size_t)info(2));
  }

  ScalarType *val_dev = nullptr;
  int *column_index_dev = nullptr;
  int *row_ptr_dev = nullptr;

private:
  void release()
  {
    if (_csr)
      mfmgb_csr_destroy(_handle->ctx, _csr);
    _csr = nullptr;
  }
  int64_t info(int which) const
  {
    int64_t v[3] = {0, 0, 0};
    if (_csr)
      mfmgb_csr_info(_csr, &v[0], &v[1], &v[2]);
    return v[which];
  }
  std::shared_ptr<CudaHandle const> _handle;
  mfmgb_csr *_csr = nullptr;
  MPI_Comm _comm = MPI_COMM_SELF;
};

namespace internal
{
// C = A B on the host (Gustavson, rows of C sorted by column).  SETUP operation: like the reference's parallel path,
// which multiplies through Trilinos on the host (include/mfmg/cuda/sparse_matrix_device.templates.cuh:417-433), the
// Galerkin product R A R^T is formed once on the host and uploaded.
inline void host_multiply_arrays(SparseMatrixDevice<double> const &a, SparseMatrixDevice<double> const &b,
                                 std::vector<int64_t> &crp, std::vector<int> &cc, std::vector<double> &cv)
{
  ASSERT_THROW(a.n() == b.m(), "multiply: inner dimensions differ");
  std::vector<int64_t> arp, brp;
  std::vector<int> ac, bc;
  std::vector<double> av, bv;
  a.copy_to_host(arp, ac, av);
  b.copy_to_host(brp, bc, bv);
  unsigned int const m = a.m(), n = b.n();
  crp.assign(1, 0);
  cc.clear();
  cv.clear();
  std::vector<double> acc(n, 0.);
  std::vector<char> used(n, 0);
  std::vector<int> cols;
  for (unsigned int i = 0; i < m; ++i)
  {
    cols.clear();
    for (int64_t ka = arp[i]; ka < arp[i + 1]; ++ka)
    {
      int const k = ac[(std::size_t)ka];
      for (int64_t kb = brp[k]; kb < brp[k + 1]; ++kb)
      {
        int const j = bc[(std::size_t)kb];
        if (!used[j])
        {
          used[j] = 1;
          cols.push_back(j);
        }
        acc[j] += av[(std::size_t)ka] * bv[(std::size_t)kb];
      }
    }
    std::sort(cols.begin(), cols.end());
    for (int j : cols)
    {
      cc.push_back(j);
      cv.push_back(acc[j]);
      acc[j] = 0.;
      used[j] = 0;
    }
    crp.push_back((int64_t)cc.size());
  }
}

inline std::shared_ptr<SparseMatrixDevice<double>> host_multiply(SparseMatrixDevice<double> const &a,
                                                                 SparseMatrixDevice<double> const &b)
{
  std::vector<int64_t> crp;
  std::vector<int> cc;
  std::vector<double> cv;
  host_multiply_arrays(a, b, crp, cc, cv);
  return std::make_shared<SparseMatrixDevice<double>>(a.handle(), a.m(), b.n(), crp, cc, cv);
}
} // namespace internal

template <typename ScalarType>
void SparseMatrixDevice<ScalarType>::mmult(SparseMatrixDevice<ScalarType> &C,
                                           SparseMatrixDevice<ScalarType> const &B) const
{
  std::vector<int64_t> crp;
  std::vector<int> cc;
  std::vector<double> cv;
  internal::host_multiply_arrays(*this, B, crp, cc, cv);
  C.reinit(_handle, m(), B.n(), crp, cc, cv);
}

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
  // Setup operations (cuda_matrix_operator.cu:132-225): the reference computes products with cuSPARSE csrgemm (serial)
  // or on the host through Trilinos (parallel).  Setup stays on the host path: host product, uploaded once.
  std::shared_ptr<Operator<vector_type>> multiply(std::shared_ptr<Operator<vector_type> const> b) const override
  {
    auto rhs = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type> const>(b);
    ASSERT_THROW(rhs != nullptr, "CudaMatrixOperator::multiply needs a CudaMatrixOperator");
    return std::make_shared<CudaMatrixOperator<vector_type>>(internal::host_multiply(*_matrix, *rhs->get_matrix()));
  }
  std::shared_ptr<Operator<vector_type>>
  multiply_transpose(std::shared_ptr<Operator<vector_type> const> b) const override
  {
    auto rhs = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type> const>(b);
    ASSERT_THROW(rhs != nullptr, "CudaMatrixOperator::multiply_transpose needs a CudaMatrixOperator");
    auto bt = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type>>(rhs->transpose());
    return std::make_shared<CudaMatrixOperator<vector_type>>(internal::host_multiply(*_matrix, *bt->get_matrix()));
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

// ---------------------------------------------------------------------------------------------------------------
// Mesh evaluators and the matrix-free operator slot
//   include/mfmg/common/mesh_evaluator.hpp, include/mfmg/cuda/cuda_mesh_evaluator.cuh:25-74,
//   include/mfmg/cuda/cuda_matrix_free_mesh_evaluator.cuh:25-104, include/mfmg/cuda/cuda_matrix_free_operator.cuh:22-77
// The reference evaluators carry a dealii::DoFHandler and AffineConstraints; without deal.II they carry the CudaHandle
// only, and -- because SETUP STAYS ON THE HOST PATH -- one extra hook, build_restrictor_matrix, through which the
// restriction matrix produced by mfmg's own AMGe setup is handed over (the reference calls AMGe_device there,
// source/cuda/cuda_hierarchy_helpers.cu:56-81).
// ---------------------------------------------------------------------------------------------------------------
class MeshEvaluator
{
public:
  virtual ~MeshEvaluator() = default;
  virtual int get_dim() const = 0;
  virtual std::string get_mesh_evaluator_type() const = 0;
};

template <int dim>
class CudaMeshEvaluator : public MeshEvaluator
{
public:
  explicit CudaMeshEvaluator(std::shared_ptr<CudaHandle const> cuda_handle) : _cuda_handle(std::move(cuda_handle)) {}
  int get_dim() const final { return dim; }
  std::string get_mesh_evaluator_type() const override { return "CudaMeshEvaluator"; }
  // user hook: fill the global system matrix (cuda_mesh_evaluator.cuh:44-49)
  virtual void evaluate_global(SparseMatrixDevice<double> &) const { throw NotImplementedExc(); }
  // user hook: local matrix of an agglomerate (cuda_mesh_evaluator.cuh:37-42) -- only the host setup needs it
  virtual void evaluate_agglomerate(SparseMatrixDevice<double> &) const { throw NotImplementedExc(); }
  // hand-over of the host setup's restriction matrix (n_coarse x n_fine), see the banner above
  virtual void build_restrictor_matrix(std::shared_ptr<ParameterTree const>, SparseMatrixDevice<double> &) const
  {
    throw NotImplementedExc("build_restrictor_matrix: bind mfmg's host AMGe setup here");
  }
  std::shared_ptr<CudaHandle const> const &get_cuda_handle() const { return _cuda_handle; }

protected:
  std::shared_ptr<CudaHandle const> _cuda_handle;
};

// dealii::DiagonalMatrix<VectorType> stand-in: the inverse diagonal a matrix-free smoother uses
class DiagonalMatrix
{
public:
  explicit DiagonalMatrix(std::shared_ptr<DeviceVector> diagonal) : _diagonal(std::move(diagonal)) {}
  DeviceVector const &get_vector() const { return *_diagonal; }

private:
  std::shared_ptr<DeviceVector> _diagonal;
};

template <int dim>
class CudaMatrixFreeMeshEvaluator : public CudaMeshEvaluator<dim>
{
public:
  using size_type = unsigned int;
  static int constexpr _dim = dim;
  explicit CudaMatrixFreeMeshEvaluator(std::shared_ptr<CudaHandle const> cuda_handle)
      : CudaMeshEvaluator<dim>(std::move(cuda_handle))
  {
  }
  std::string get_mesh_evaluator_type() const final { return "CudaMatrixFreeMeshEvaluator"; }
  // the hooks of cuda_matrix_free_mesh_evaluator.cuh:47-97 (all NotImplemented in the reference)
  virtual std::shared_ptr<DeviceVector> build_range_vector() const { throw NotImplementedExc(); }
  virtual void matrix_free_evaluate_global(DeviceVector const & /*src*/, DeviceVector & /*dst*/) const
  {
    throw NotImplementedExc();
  }
  virtual std::shared_ptr<DiagonalMatrix> matrix_free_get_diagonal_inverse() const { throw NotImplementedExc(); }
  virtual std::shared_ptr<DeviceVector> get_diagonal() const { throw NotImplementedExc(); }
  // non-null when the evaluator is served by the library's own operator: the fused V-cycle then uses it directly
  virtual mfmgb_mf const *native_operator() const { return nullptr; }
};

// The library's evaluator for the reference's matrix-free test problem (tests/laplace_matrix_free.hpp:30-199) on a
// uniform Cartesian grid with lexicographic DoFs: fills every hook above with the hand-written sm_100a kernels.
template <int dim>
class LaplaceMatrixFreeMeshEvaluator : public CudaMatrixFreeMeshEvaluator<dim>
{
public:
  // coef: [n_cells][(degree+1)^dim] coefficient at the quadrature points; constrained: [n_dofs] Dirichlet flags
  LaplaceMatrixFreeMeshEvaluator(std::shared_ptr<CudaHandle const> cuda_handle, int degree, std::vector<int64_t> const &cells,
                                 std::vector<double> const &h, std::vector<double> const &coef,
                                 std::vector<uint8_t> const &constrained)
      : CudaMatrixFreeMeshEvaluator<dim>(std::move(cuda_handle))
  {
    ASSERT_THROW((int)cells.size() == dim && (int)h.size() == dim, "LaplaceMatrixFreeMeshEvaluator: bad grid description");
    check_status(this->_cuda_handle->ctx, mfmgb_mf_laplace_create(this->_cuda_handle->ctx, dim, degree, cells.data(), h.data(),
                                                                  coef.data(), constrained.data(), &_mf));
  }
  ~LaplaceMatrixFreeMeshEvaluator() override { mfmgb_mf_destroy(this->_cuda_handle->ctx, _mf); }
  std::shared_ptr<DeviceVector> build_range_vector() const override
  {
    return std::make_shared<DeviceVector>(this->_cuda_handle, (std::size_t)mfmgb_mf_size(_mf));
  }
  void matrix_free_evaluate_global(DeviceVector const &src, DeviceVector &dst) const override
  {
    check_status(this->_cuda_handle->ctx, mfmgb_mf_apply(this->_cuda_handle->ctx, _mf, src.get_values(), dst.get_values()));
  }
  std::shared_ptr<DeviceVector> get_diagonal() const override
  {
    auto d = build_range_vector();
    check_status(this->_cuda_handle->ctx, mfmgb_mf_diagonal(this->_cuda_handle->ctx, _mf, d->get_values()));
    return d;
  }
  std::shared_ptr<DiagonalMatrix> matrix_free_get_diagonal_inverse() const override
  {
    auto d = get_diagonal();
    auto h = d->export_to_host();
    for (auto &v : h)
      v = 1. / v;
    d->import_from_host(h);
    return std::make_shared<DiagonalMatrix>(d);
  }
  mfmgb_mf const *native_operator() const override { return _mf; }

private:
  mfmgb_mf *_mf = nullptr;
};

// include/mfmg/cuda/cuda_matrix_free_operator.cuh:22-77, source/cuda/cuda_matrix_free_operator.cu:24-140
template <int dim, typename VectorType>
class CudaMatrixFreeOperator final : public Operator<VectorType>
{
public:
  using vector_type = VectorType;
  using size_type = std::size_t;
  explicit CudaMatrixFreeOperator(std::shared_ptr<CudaMatrixFreeMeshEvaluator<dim>> matrix_free_mesh_evaluator)
      : _mesh_evaluator(std::move(matrix_free_mesh_evaluator))
  {
  }
  void vmult(vector_type &dst, vector_type const &src) const { _mesh_evaluator->matrix_free_evaluate_global(src, dst); }
  void apply(vector_type const &x, vector_type &y, OperatorMode mode = OperatorMode::NO_TRANS) const override
  {
    if (mode != OperatorMode::NO_TRANS)
      throw NotImplementedExc(); // cuda_matrix_free_operator.cu:62-66
    vmult(y, x);
  }
  std::shared_ptr<Operator<vector_type>> transpose() const override { throw NotImplementedExc(); }
  std::shared_ptr<Operator<vector_type>> multiply(std::shared_ptr<Operator<vector_type> const>) const override
  {
    throw NotImplementedExc();
  }
  // this * b^T column by column: the operator applied to every row of b (the restrictor), a SETUP operation like
  // DealIIMatrixFreeOperator::multiply_transpose on the reference's host path; the result is assembled
  std::shared_ptr<Operator<vector_type>> multiply_transpose(std::shared_ptr<Operator<vector_type> const> b) const override
  {
    auto r = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type> const>(b);
    ASSERT_THROW(r != nullptr, "CudaMatrixFreeOperator::multiply_transpose needs a CudaMatrixOperator");
    auto rm = r->get_matrix();
    std::vector<int64_t> rp;
    std::vector<int> rc;
    std::vector<double> rv;
    rm->copy_to_host(rp, rc, rv);
    unsigned int const nc = rm->m(), n = rm->n();
    auto v = build_domain_vector(), w = build_range_vector();
    std::vector<std::vector<std::pair<int, double>>> rows(n); // (A R^T) by row: entry (i, c)
    std::vector<double> dense(n, 0.);
    for (unsigned int c = 0; c < nc; ++c)
    {
      std::fill(dense.begin(), dense.end(), 0.);
      for (int64_t k = rp[c]; k < rp[c + 1]; ++k)
        dense[(std::size_t)rc[(std::size_t)k]] = rv[(std::size_t)k];
      v->import_from_host(dense);
      vmult(*w, *v);
      auto col = w->export_to_host();
      for (unsigned int i = 0; i < n; ++i)
        if (col[i] != 0.)
          rows[i].emplace_back((int)c, col[i]);
    }
    std::vector<int64_t> orp(1, 0);
    std::vector<int> oc;
    std::vector<double> ov;
    for (unsigned int i = 0; i < n; ++i)
    {
      for (auto const &e : rows[i])
      {
        oc.push_back(e.first);
        ov.push_back(e.second);
      }
      orp.push_back((int64_t)oc.size());
    }
    return std::make_shared<CudaMatrixOperator<vector_type>>(
        std::make_shared<SparseMatrixDevice<double>>(_mesh_evaluator->get_cuda_handle(), n, nc, orp, oc, ov));
  }
  std::shared_ptr<vector_type> build_domain_vector() const override { return _mesh_evaluator->build_range_vector(); }
  std::shared_ptr<vector_type> build_range_vector() const override { return _mesh_evaluator->build_range_vector(); }
  size_type grid_complexity() const override { return build_range_vector()->size(); }
  size_type operator_complexity() const override { throw NotImplementedExc(); } // cuda_matrix_free_operator.cu:128-133
  std::shared_ptr<DiagonalMatrix> get_diagonal_inverse() const { return _mesh_evaluator->matrix_free_get_diagonal_inverse(); }
  std::shared_ptr<CudaHandle const> const &get_cuda_handle() const { return _mesh_evaluator->get_cuda_handle(); }
  std::shared_ptr<CudaMatrixFreeMeshEvaluator<dim>> const &get_mesh_evaluator() const { return _mesh_evaluator; }

private:
  std::shared_ptr<CudaMatrixFreeMeshEvaluator<dim>> _mesh_evaluator;
};

// Jacobi smoother on a matrix-free operator: x <- x - D^-1 (A x - b) from the evaluator's inverse diagonal (the
// composition the reference's host path uses for matrix-free levels, source/dealii/dealii_matrix_free_smoother.cc:67-79,
// with PreconditionChebyshev of degree 0 and theta = 1).  Unfused: the fused form lives in Hierarchy::apply.
template <int dim, typename VectorType>
class CudaMatrixFreeSmoother : public Smoother<VectorType>
{
public:
  using vector_type = VectorType;
  CudaMatrixFreeSmoother(std::shared_ptr<Operator<vector_type> const> op, std::shared_ptr<ParameterTree const> params)
      : Smoother<vector_type>(op, params)
  {
    auto mf = std::dynamic_pointer_cast<CudaMatrixFreeOperator<dim, vector_type> const>(this->_operator);
    ASSERT_THROW(mf != nullptr, "CudaMatrixFreeSmoother must be constructed from a CudaMatrixFreeOperator");
    _inverse_diagonal = mf->get_diagonal_inverse();
    auto diag = mf->get_mesh_evaluator()->get_diagonal();
    auto const &ctx = diag->handle()->ctx;
    check_status(ctx, mfmgb_jacobi_setup_diag(ctx, diag->get_values(), (int64_t)diag->size(), 1.0, &_jacobi));
    check_status(ctx, mfmgb_ctx_synchronize(ctx));
  }
  ~CudaMatrixFreeSmoother() override { mfmgb_jacobi_destroy(_inverse_diagonal->get_vector().handle()->ctx, _jacobi); }
  void apply(vector_type const &b, vector_type &x) const override
  {
    vector_type r(b);
    this->_operator->apply(x, r); // r = A x
    r.add(-1., b);                // r = A x - b
    auto const &ctx = _inverse_diagonal->get_vector().handle()->ctx;
    check_status(ctx, mfmgb_jacobi_apply_residual(ctx, _jacobi, r.get_values(), x.get_values())); // x -= D^-1 r
  }

private:
  std::shared_ptr<DiagonalMatrix> _inverse_diagonal;
  mfmgb_jacobi *_jacobi = nullptr;
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

// dealii::TimerOutput stand-in (hierarchy.hpp:36-47): wall-clock seconds per named section
class TimerOutput
{
public:
  void enter_subsection(std::string const &section)
  {
    _open.emplace_back(section, std::chrono::steady_clock::now());
  }
  void leave_subsection()
  {
    if (_open.empty())
      return;
    auto const dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - _open.back().second).count();
    _seconds[_open.back().first] += dt;
    _open.pop_back();
  }
  std::map<std::string, double> const &get_summary_data() const { return _seconds; }

private:
  std::vector<std::pair<std::string, std::chrono::steady_clock::time_point>> _open;
  std::map<std::string, double> _seconds;
};

inline void timer_enter_subsection(std::shared_ptr<TimerOutput> const &timer, std::string const &section)
{
  if (timer)
    timer->enter_subsection(section);
}
inline void timer_leave_subsection(std::shared_ptr<TimerOutput> const &timer)
{
  if (timer)
    timer->leave_subsection();
}

// include/mfmg/common/hierarchy_helpers.hpp:27-62
template <typename VectorType>
class HierarchyHelpers
{
public:
  using vector_type = VectorType;
  virtual ~HierarchyHelpers() = default;
  virtual std::shared_ptr<Operator<vector_type>> get_global_operator(std::shared_ptr<MeshEvaluator> mesh_evaluator) = 0;
  virtual std::shared_ptr<Operator<vector_type>> build_restrictor(MPI_Comm comm, std::shared_ptr<MeshEvaluator> mesh_evaluator,
                                                                  std::shared_ptr<ParameterTree const> params) = 0;
  virtual std::shared_ptr<Smoother<vector_type>> build_smoother(std::shared_ptr<Operator<vector_type> const> op,
                                                                std::shared_ptr<ParameterTree const> params) = 0;
  virtual std::shared_ptr<Solver<vector_type>> build_coarse_solver(std::shared_ptr<Operator<vector_type> const> op,
                                                                   std::shared_ptr<ParameterTree const> params) = 0;
  // hierarchy_helpers.hpp:52-57 (the matrix-free helpers override it on the host path)
  virtual std::shared_ptr<Operator<vector_type>>
  fast_multiply_transpose(std::shared_ptr<Operator<vector_type> const> a, std::shared_ptr<Operator<vector_type> const> r)
  {
    return a->multiply_transpose(r);
  }
};

// include/mfmg/cuda/cuda_hierarchy_helpers.cuh:24-54, source/cuda/cuda_hierarchy_helpers.cu:29-121
template <int dim, typename VectorType>
class CudaHierarchyHelpers : public HierarchyHelpers<VectorType>
{
public:
  using vector_type = VectorType;
  explicit CudaHierarchyHelpers(std::shared_ptr<CudaHandle const> cuda_handle) : _cuda_handle(std::move(cuda_handle)) {}
  std::shared_ptr<Operator<vector_type>> get_global_operator(std::shared_ptr<MeshEvaluator> mesh_evaluator) override
  {
    if (_operator == nullptr)
    {
      auto cuda_mesh_evaluator = std::dynamic_pointer_cast<CudaMeshEvaluator<dim>>(mesh_evaluator);
      ASSERT_THROW(cuda_mesh_evaluator != nullptr, "CudaHierarchyHelpers needs a CudaMeshEvaluator");
      auto mf_evaluator = std::dynamic_pointer_cast<CudaMatrixFreeMeshEvaluator<dim>>(mesh_evaluator);
      if (mf_evaluator)
        _operator.reset(new CudaMatrixFreeOperator<dim, vector_type>(mf_evaluator));
      else
      {
        auto system_matrix = std::make_shared<SparseMatrixDevice<double>>();
        cuda_mesh_evaluator->evaluate_global(*system_matrix); // user function fills the system matrix
        _operator.reset(new CudaMatrixOperator<vector_type>(system_matrix));
      }
    }
    return _operator;
  }
  std::shared_ptr<Operator<vector_type>> build_restrictor(MPI_Comm, std::shared_ptr<MeshEvaluator> mesh_evaluator,
                                                          std::shared_ptr<ParameterTree const> params) override
  {
    auto cuda_mesh_evaluator = std::dynamic_pointer_cast<CudaMeshEvaluator<dim>>(mesh_evaluator);
    ASSERT_THROW(cuda_mesh_evaluator != nullptr, "CudaHierarchyHelpers needs a CudaMeshEvaluator");
    auto restrictor_matrix = std::make_shared<SparseMatrixDevice<double>>();
    cuda_mesh_evaluator->build_restrictor_matrix(params, *restrictor_matrix); // the host AMGe setup's result
    return std::make_shared<CudaMatrixOperator<vector_type>>(restrictor_matrix);
  }
  std::shared_ptr<Smoother<vector_type>> build_smoother(std::shared_ptr<Operator<vector_type> const> op,
                                                        std::shared_ptr<ParameterTree const> params) override
  {
    if (std::dynamic_pointer_cast<CudaMatrixFreeOperator<dim, vector_type> const>(op))
      return std::make_shared<CudaMatrixFreeSmoother<dim, vector_type>>(op, params);
    return std::make_shared<CudaSmoother<vector_type>>(op, params);
  }
  std::shared_ptr<Solver<vector_type>> build_coarse_solver(std::shared_ptr<Operator<vector_type> const> op,
                                                           std::shared_ptr<ParameterTree const> params) override
  {
    return std::make_shared<CudaSolver<vector_type>>(*_cuda_handle, op, params);
  }

private:
  std::shared_ptr<CudaHandle const> _cuda_handle;
  std::shared_ptr<Operator<vector_type>> _operator;
};

// include/mfmg/common/hierarchy.hpp:49-153: dispatch on the evaluator's type string and dimension
template <typename VectorType>
std::unique_ptr<HierarchyHelpers<VectorType>> create_hierarchy_helpers(std::shared_ptr<MeshEvaluator const> evaluator)
{
  std::unique_ptr<HierarchyHelpers<VectorType>> hierarchy_helpers;
  std::string const evaluator_type = evaluator->get_mesh_evaluator_type();
  if (evaluator_type == "CudaMeshEvaluator" || evaluator_type == "CudaMatrixFreeMeshEvaluator")
  {
    int const dim = evaluator->get_dim();
    if (dim == 2)
      hierarchy_helpers.reset(new CudaHierarchyHelpers<2, VectorType>(
          std::dynamic_pointer_cast<CudaMeshEvaluator<2> const>(evaluator)->get_cuda_handle()));
    else if (dim == 3)
      hierarchy_helpers.reset(new CudaHierarchyHelpers<3, VectorType>(
          std::dynamic_pointer_cast<CudaMeshEvaluator<3> const>(evaluator)->get_cuda_handle()));
    else
      throw NotImplementedExc();
  }
  else // the DealII* evaluators belong to the reference's host path
    throw NotImplementedExc("create_hierarchy_helpers: " + evaluator_type + " is not a device evaluator");
  return hierarchy_helpers;
}

// include/mfmg/common/hierarchy.hpp:155-315
template <typename VectorType>
class Hierarchy
{
public:
  using vector_type = VectorType;

  // The reference constructor (hierarchy.hpp:159-236): setup from a MeshEvaluator, statement by statement.  The
  // operators come from the evaluator's hooks (setup stays on the host path); smoothers, the coarse solver and the
  // Galerkin product are built exactly where the reference builds them.
  Hierarchy(MPI_Comm comm, std::shared_ptr<MeshEvaluator> evaluator, std::shared_ptr<ParameterTree> params = nullptr,
            std::shared_ptr<TimerOutput> timer = nullptr)
      : _timer(std::move(timer))
  {
    timer_enter_subsection(_timer, "Setup");
    if (!params)
      params = std::make_shared<ParameterTree>();
    auto hierarchy_helpers = create_hierarchy_helpers<vector_type>(evaluator);
    _is_preconditioner = params->get("is preconditioner", true);        // :168
    _n_smoothing_steps = params->get("smoother.n_smoothing_steps", 1u); // :169
    unsigned int const num_levels = params->get("max levels", 2u);      // :171
    ASSERT_THROW(num_levels >= 1 && num_levels <= 2, "max levels > 2 needs one evaluator per level (hierarchy.hpp:209-210 "
                                                     "re-uses the fine evaluator, which the reference itself marks broken)");
    _levels.resize(num_levels);
    for (unsigned int level_index = 0; level_index < num_levels; ++level_index)
    {
      auto &level_fine = _levels[level_index];
      if (level_index == 0)
      {
        timer_enter_subsection(_timer, "Setup: build global operator");
        level_fine.set_operator(hierarchy_helpers->get_global_operator(evaluator)); // :178-180
        timer_leave_subsection(_timer);
      }
      if (level_index == num_levels - 1)
      {
        timer_enter_subsection(_timer, "Setup: build coarse solver");
        level_fine.set_solver(hierarchy_helpers->build_coarse_solver(level_fine.get_operator(), params)); // :192-195
        timer_leave_subsection(_timer);
        break;
      }
      timer_enter_subsection(_timer, "Setup: build smoother");
      level_fine.set_smoother(hierarchy_helpers->build_smoother(level_fine.get_operator(), params)); // :203-205
      timer_leave_subsection(_timer);
      auto &level_coarse = _levels[level_index + 1];
      timer_enter_subsection(_timer, "Setup: build restrictor");
      auto restrictor = hierarchy_helpers->build_restrictor(comm, evaluator, params); // :212-214
      level_coarse.set_restrictor(restrictor);
      timer_leave_subsection(_timer);
      timer_enter_subsection(_timer, "Setup: build coarse operator");
      auto a = level_fine.get_operator();
      auto ap = hierarchy_helpers->fast_multiply_transpose(a, restrictor); // :222-226  A R^T
      level_coarse.set_operator(restrictor->multiply(ap));                 // :229-231  R (A R^T)
      timer_leave_subsection(_timer);
    }
    build_fused(params);
    timer_leave_subsection(_timer);
  }

  // Setup already done on the host path: the level operators A_l and restrictors R_l are given; smoothers and the coarse
  // solver are built exactly as hierarchy.hpp:183-234 does.
  Hierarchy(std::shared_ptr<CudaHandle const> handle,
            std::vector<std::shared_ptr<CudaMatrixOperator<vector_type>>> const &operators,
            std::vector<std::shared_ptr<CudaMatrixOperator<vector_type>>> const &restrictors,
            std::shared_ptr<ParameterTree> params = nullptr)
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
        _levels[l].set_solver(std::make_shared<CudaSolver<vector_type>>(*handle, operators[l], params)); // :194
    }
    build_fused(params);
    ASSERT_THROW(_fused != nullptr, "Hierarchy: the fused device path could not be built");
  }
  ~Hierarchy()
  {
    if (_fused)
      mfmgb_hierarchy_destroy(_ctx, _fused);
  }
  Hierarchy(Hierarchy const &) = delete;

  // hierarchy.hpp:238-244
  void vmult(vector_type &x, vector_type const &b) const { apply(b, x, 0); }

  // hierarchy.hpp:246-309 -- fused kernels, preallocated workspaces, one launch sequence.  Hierarchies whose level
  // operators are user objects the library does not know run the same algorithm through the abstract interfaces.
  void apply(vector_type const &b, vector_type &x, int level_index = 0) const
  {
    timer_enter_subsection(_timer, "Apply");
    if (_fused)
      check_status(_ctx, mfmgb_hierarchy_apply(_ctx, _fused, b.get_values(), x.get_values(), level_index));
    else
      apply_generic(b, x, level_index);
    timer_leave_subsection(_timer);
  }
  bool is_fused() const { return _fused != nullptr; }

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

  // hierarchy.hpp:311-312 (build_range_vector of the finest operator)
  std::shared_ptr<vector_type> build_range_vector() const { return _levels[0].get_operator()->build_range_vector(); }

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
  // the fused device V-cycle needs level operators the library owns: CSR matrices, or (level 0) its matrix-free operator
  template <int dim>
  static mfmgb_mf const *native_mf(std::shared_ptr<Operator<vector_type> const> const &op, mfmgb_ctx *&ctx)
  {
    auto mf = std::dynamic_pointer_cast<CudaMatrixFreeOperator<dim, vector_type> const>(op);
    if (!mf)
      return nullptr;
    ctx = mf->get_cuda_handle()->ctx;
    return mf->get_mesh_evaluator()->native_operator();
  }
  void build_fused(std::shared_ptr<ParameterTree> const &params)
  {
    unsigned int const num_levels = (unsigned int)_levels.size();
    std::vector<mfmgb_csr const *> ops(num_levels, nullptr), res(num_levels, nullptr);
    mfmgb_mf const *mf0 = nullptr;
    for (unsigned int l = 0; l < num_levels; ++l)
    {
      auto op = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type> const>(_levels[l].get_operator());
      if (op)
      {
        ops[l] = op->get_matrix()->c_handle();
        _ctx = op->get_matrix()->handle()->ctx;
      }
      else if (l == 0 && ((mf0 = native_mf<3>(_levels[0].get_operator(), _ctx)) != nullptr ||
                          (mf0 = native_mf<2>(_levels[0].get_operator(), _ctx)) != nullptr))
        ;
      else
        return; // a user operator: the abstract composition serves the hierarchy
      if (l > 0)
      {
        auto r = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type> const>(_levels[l].get_restrictor());
        if (!r)
          return;
        res[l] = r->get_matrix()->c_handle();
      }
    }
    std::string smoother = params->get("smoother.type", "Jacobi");
    std::transform(smoother.begin(), smoother.end(), smoother.begin(), ::tolower);
    check_status(_ctx, mfmgb_hierarchy_create(_ctx, (int)num_levels, (int)_n_smoothing_steps, _is_preconditioner ? 1 : 0, 1.0,
                                              &_fused));
    for (unsigned int l = 0; l < num_levels; ++l)
    {
      if (ops[l])
        check_status(_ctx, mfmgb_hierarchy_set_operator(_fused, (int)l, ops[l]));
      else
        check_status(_ctx, mfmgb_hierarchy_set_mf_operator(_fused, mf0));
      if (l > 0)
        check_status(_ctx, mfmgb_hierarchy_set_restrictor(_fused, (int)l, res[l], nullptr));
    }
    if (smoother == "chebyshev") // source/dealii/dealii_matrix_free_smoother.cc:34-60
      check_status(_ctx, mfmgb_hierarchy_set_smoother_chebyshev(_fused, params->get("smoother.degree", 0),
                                                                params->get("smoother.smoothing_range", 0.),
                                                                params->get("smoother.max_eigenvalue", 1.), 8));
    check_status(_ctx, mfmgb_hierarchy_finalize(_ctx, _fused));
  }

  std::shared_ptr<TimerOutput> _timer;
  mfmgb_ctx *_ctx = nullptr;
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

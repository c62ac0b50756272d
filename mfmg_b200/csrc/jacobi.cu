// jacobi.cu -- Jacobi smoother: D^-1 extraction and the fused sweep.
// Replaces mfmg::CudaSmoother (source/cuda/cuda_smoother.cu): the reference stores D^-1 as a
// CSR matrix and applies it with a second cuSPARSE SpMV, two vector copies and two axpys; here
// the whole sweep x <- x - omega D^-1 (A x - b) is the epilogue of one SpMV launch.
#include "jacobi.cuh"
#include "vecops.cuh"

using namespace mfmgb;

namespace
{
constexpr int kBlock = 256;

// one thread per row scans its row for the diagonal (extract_inv_diag, cuda_smoother.cu:86-96,
// but indexed by the LOCAL row so it is also right on ranks != 0)
template <typename OffT>
__global__ void __launch_bounds__(kBlock)
    inv_diag_kernel(int64_t n, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                    const double *__restrict__ val, double *__restrict__ dinv, int *__restrict__ missing)
{
  const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (row >= n)
    return;
  double d = 0.;
  bool found = false;
  for (OffT k = rowptr[row]; k < rowptr[row + 1]; ++k)
    if (col[k] == row)
    {
      d = val[k];
      found = true;
    }
  dinv[row] = found ? 1. / d : 0.;
  if (!found || d == 0.)
    atomicAdd(missing, 1);
}

__global__ void __launch_bounds__(kBlock)
    invert_kernel(int64_t n, const double *__restrict__ diag, double *__restrict__ dinv)
{
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n)
    dinv[i] = 1. / diag[i];
}

// x = 0 - omega * dinv * (A*0 - b): bitwise what the generic sweep gives for x == 0
__device__ __forceinline__ double zero_guess_value(double dinv, double b, double omega)
{
  const double r = __dsub_rn(0., b);
  double t = __dmul_rn(dinv, r);
  if (omega != 1.)
    t = __dmul_rn(omega, t);
  return __dsub_rn(0., t);
}

__global__ void __launch_bounds__(kBlock)
    zero_guess_kernel(int64_t n, const double *__restrict__ dinv, const double *__restrict__ b, double omega,
                      double *__restrict__ x)
{
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n)
    x[i] = zero_guess_value(dinv[i], b[i], omega);
}

// x <- x - omega D^-1 r with r = A x - b formed by the caller (operators that are not CSR matrices)
__global__ void __launch_bounds__(kBlock)
    residual_update_kernel(int64_t n, const double *__restrict__ dinv, const double *__restrict__ r, double omega,
                           double *__restrict__ x)
{
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n)
  {
    double t = __dmul_rn(dinv[i], r[i]);
    if (omega != 1.)
      t = __dmul_rn(omega, t);
    x[i] = __dsub_rn(x[i], t);
  }
}

// 16-byte aligned vectors: two entries per thread and load (24 B of traffic per entry, nothing else to hide latency)
__global__ void __launch_bounds__(kBlock)
    zero_guess_kernel_v2(int64_t n, const double *__restrict__ dinv, const double *__restrict__ b, double omega,
                         double *__restrict__ x)
{
  const int64_t i = 2 * ((int64_t)blockIdx.x * kBlock + threadIdx.x);
  if (i + 1 < n)
  {
    const double2 d = *reinterpret_cast<const double2 *>(dinv + i), bb = *reinterpret_cast<const double2 *>(b + i);
    *reinterpret_cast<double2 *>(x + i) =
        make_double2(zero_guess_value(d.x, bb.x, omega), zero_guess_value(d.y, bb.y, omega));
  }
  else if (i < n)
    x[i] = zero_guess_value(dinv[i], b[i], omega);
}
} // namespace

extern "C"
{
  MFMGB_API int mfmgb_jacobi_setup(mfmgb_ctx *ctx, const mfmgb_csr *A, double omega, mfmgb_jacobi **out)
  {
    MFMGB_REQUIRE(ctx, ctx && A && out, "mfmgb_jacobi_setup: bad arguments");
    // The matrix must be square (ASSERT at cuda_smoother.cu:118-121); local partitions may have
    // extra ghost columns, hence >=.
    MFMGB_REQUIRE(ctx, A->n_cols >= A->n_rows, "mfmgb_jacobi_setup: the matrix is not square");
    *out = nullptr;
    mfmgb_jacobi *J = new mfmgb_jacobi();
    J->n = A->n_rows;
    J->n_vec = A->n_cols; // a row-partitioned operator gathers from [owned | ghost]: the copy keeps the ghost tail
    J->omega = omega;
    MFMGB_CUDA(ctx, cudaMalloc(&J->dinv, sizeof(double) * (size_t)(J->n + 2)));
    MFMGB_CUDA(ctx, cudaMalloc(&J->tmp, sizeof(double) * (size_t)(J->n_vec + 2)));
    int *missing_dev = nullptr;
    MFMGB_CUDA(ctx, cudaMalloc(&missing_dev, sizeof(int)));
    MFMGB_CUDA(ctx, cudaMemsetAsync(missing_dev, 0, sizeof(int), ctx->stream));
    if (J->n > 0)
    {
      const unsigned nb = (unsigned)ceil_div(J->n, kBlock);
      if (A->off64)
        inv_diag_kernel<int64_t><<<nb, kBlock, 0, ctx->stream>>>(J->n, (const int64_t *)A->rowptr, A->col, A->val,
                                                                 J->dinv, missing_dev);
      else
        inv_diag_kernel<int32_t><<<nb, kBlock, 0, ctx->stream>>>(J->n, (const int32_t *)A->rowptr, A->col, A->val,
                                                                 J->dinv, missing_dev);
      MFMGB_LAUNCHED(ctx);
    }
    int missing = 0;
    MFMGB_CUDA(ctx, cudaMemcpyAsync(&missing, missing_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(missing_dev);
    if (missing)
    {
      cudaFree(J->dinv);
      cudaFree(J->tmp);
      delete J;
      return fail(ctx, MFMGB_ERR_SINGULAR, "mfmgb_jacobi_setup: %d rows have no (or a zero) diagonal entry", missing);
    }
    *out = J;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_jacobi_setup_diag(mfmgb_ctx *ctx, const double *diag_dev, int64_t n, double omega,
                                        mfmgb_jacobi **out)
  {
    MFMGB_REQUIRE(ctx, ctx && out && n >= 0 && (n == 0 || diag_dev), "mfmgb_jacobi_setup_diag: bad arguments");
    mfmgb_jacobi *J = new mfmgb_jacobi();
    J->n = n;
    J->n_vec = n;
    J->omega = omega;
    MFMGB_CUDA(ctx, cudaMalloc(&J->dinv, sizeof(double) * (size_t)(n + 2)));
    MFMGB_CUDA(ctx, cudaMalloc(&J->tmp, sizeof(double) * (size_t)(n + 2)));
    if (n > 0)
    {
      invert_kernel<<<(unsigned)ceil_div(n, kBlock), kBlock, 0, ctx->stream>>>(n, diag_dev, J->dinv);
      MFMGB_LAUNCHED(ctx);
    }
    *out = J;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_jacobi_destroy(mfmgb_ctx *ctx, mfmgb_jacobi *J)
  {
    if (!J)
      return MFMGB_OK;
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(J->dinv);
    cudaFree(J->tmp);
    delete J;
    return MFMGB_OK;
  }

  MFMGB_API const double *mfmgb_jacobi_inv_diag(const mfmgb_jacobi *J) { return J ? J->dinv : nullptr; }

  MFMGB_API int mfmgb_jacobi_apply_oop(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const mfmgb_csr *A, const double *b,
                                       const double *x_in, double *x_out)
  {
    MFMGB_REQUIRE(ctx, ctx && J && A && b && x_in && x_out, "mfmgb_jacobi_apply_oop: bad arguments");
    MFMGB_REQUIRE(ctx, x_in != x_out, "mfmgb_jacobi_apply_oop: x_in and x_out must not alias");
    MFMGB_REQUIRE(ctx, A->n_rows == J->n, "mfmgb_jacobi_apply_oop: size mismatch");
    EpiArgs e;
    e.y = x_out;
    e.b = b;
    e.dinv = J->dinv;
    e.xin = x_in;
    e.omega = J->omega;
    return csr_apply(ctx, A, x_in, Epi::Jacobi, e);
  }

  MFMGB_API int mfmgb_jacobi_apply(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const mfmgb_csr *A, const double *b, double *x)
  {
    MFMGB_REQUIRE(ctx, ctx && J && A && b && x, "mfmgb_jacobi_apply: bad arguments");
    // Jacobi reads neighbouring x entries while others are written: keep the old iterate (with its ghost tail when
    // the operator is a row-partitioned block, n_cols > n_rows).
    MFMGB_REQUIRE(ctx, A->n_cols <= J->n_vec, "mfmgb_jacobi_apply: the operator gathers from more columns than the "
                                              "smoother was set up for; use mfmgb_jacobi_apply_oop");
    MFMGB_CUDA(ctx, cudaMemcpyAsync(J->tmp, x, sizeof(double) * (size_t)A->n_cols, cudaMemcpyDeviceToDevice, ctx->stream));
    return mfmgb_jacobi_apply_oop(ctx, J, A, b, J->tmp, x);
  }

  MFMGB_API int mfmgb_jacobi_apply_residual(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const double *r, double *x)
  {
    MFMGB_REQUIRE(ctx, ctx && J && r && x && r != x, "mfmgb_jacobi_apply_residual: bad arguments");
    if (J->n == 0)
      return MFMGB_OK;
    residual_update_kernel<<<(unsigned)ceil_div(J->n, kBlock), kBlock, 0, ctx->stream>>>(J->n, J->dinv, r, J->omega, x);
    MFMGB_LAUNCHED(ctx);
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_jacobi_apply_zero_guess(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const double *b, double *x)
  {
    MFMGB_REQUIRE(ctx, ctx && J && b && x, "mfmgb_jacobi_apply_zero_guess: bad arguments");
    if (J->n == 0)
      return MFMGB_OK;
    if (((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(J->dinv)) & 15) == 0)
      zero_guess_kernel_v2<<<(unsigned)ceil_div(ceil_div(J->n, 2), kBlock), kBlock, 0, ctx->stream>>>(J->n, J->dinv, b,
                                                                                                     J->omega, x);
    else
      zero_guess_kernel<<<(unsigned)ceil_div(J->n, kBlock), kBlock, 0, ctx->stream>>>(J->n, J->dinv, b, J->omega, x);
    MFMGB_LAUNCHED(ctx);
    return MFMGB_OK;
  }
}

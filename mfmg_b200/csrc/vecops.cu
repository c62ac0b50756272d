// vecops.cu -- fill / axpy / deterministic dot.
#include "vecops.cuh"

namespace mfmgb
{
namespace
{
constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) fill_kernel(double *__restrict__ v, double value, int64_t n)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    v[i] = value;
}

__global__ void __launch_bounds__(kBlock)
    axpy_kernel(double *__restrict__ y, double a, const double *__restrict__ x, int64_t n)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    y[i] = y[i] + a * x[i];
}

__global__ void __launch_bounds__(kBlock)
    dot_partial_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t n,
                       double *__restrict__ partials)
{
  __shared__ double sm[kBlock / 32];
  double s = 0.;
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    s = fma(a[i], b[i], s);
  s = block_sum<kBlock>(s, sm);
  if (threadIdx.x == 0)
    partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024)
    finalize_kernel(const double *__restrict__ partials, int nblocks, int stride, double *__restrict__ result)
{
  __shared__ double sm[32];
  const double *p = partials + (size_t)blockIdx.x * stride;
  double s = 0.;
  for (int i = threadIdx.x; i < nblocks; i += 1024)
    s += p[i];
  s = block_sum<1024>(s, sm);
  if (threadIdx.x == 0)
    result[blockIdx.x] = s;
}
} // namespace

int reduce_blocks(const mfmgb_ctx *ctx, int64_t n)
{
  int64_t nb = ceil_div(n, (int64_t)kBlock * 8);
  const int64_t cap = (int64_t)ctx->num_sms * 8;
  if (nb > cap)
    nb = cap;
  if (nb > ctx->red_capacity)
    nb = ctx->red_capacity;
  if (nb < 1)
    nb = 1;
  return (int)nb;
}

int vec_fill(mfmgb_ctx *ctx, double *v, double value, int64_t n)
{
  if (n == 0)
    return MFMGB_OK;
  if (value == 0.)
  {
    MFMGB_CUDA(ctx, cudaMemsetAsync(v, 0, sizeof(double) * (size_t)n, ctx->stream));
    return MFMGB_OK;
  }
  const int nb = (int)std::min<int64_t>(ceil_div(n, kBlock), (int64_t)ctx->num_sms * 16);
  fill_kernel<<<nb, kBlock, 0, ctx->stream>>>(v, value, n);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

int vec_axpy(mfmgb_ctx *ctx, double *y, double a, const double *x, int64_t n)
{
  if (n == 0)
    return MFMGB_OK;
  const int nb = (int)std::min<int64_t>(ceil_div(n, kBlock), (int64_t)ctx->num_sms * 16);
  axpy_kernel<<<nb, kBlock, 0, ctx->stream>>>(y, a, x, n);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

int reduce_finalize(mfmgb_ctx *ctx, const double *partials, int nblocks, int stride, int n_results,
                    double *result_dev)
{
  finalize_kernel<<<n_results, 1024, 0, ctx->stream>>>(partials, nblocks, stride, result_dev);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

int vec_dot_async(mfmgb_ctx *ctx, const double *a, const double *b, int64_t n, double *result_dev)
{
  const int nb = reduce_blocks(ctx, n);
  dot_partial_kernel<<<nb, kBlock, 0, ctx->stream>>>(a, b, n, ctx->red_partials);
  MFMGB_LAUNCHED(ctx);
  return reduce_finalize(ctx, ctx->red_partials, nb, ctx->red_capacity, 1, result_dev);
}
} // namespace mfmgb

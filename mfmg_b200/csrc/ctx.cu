// ctx.cu -- context lifetime and the BLAS-1 vector operations on the path.
#include <algorithm>

#include "common.cuh"
#include "vecops.cuh"

using namespace mfmgb;

extern "C"
{
  MFMGB_API const char *mfmgb_version(void) { return "mfmg_b200 0.1.0 (sm_100a)"; }

  MFMGB_API const char *mfmgb_last_error(mfmgb_ctx *ctx)
  {
    return ctx ? ctx->error.c_str() : tls_error().c_str();
  }

  MFMGB_API int mfmgb_ctx_create(int device, void *stream, mfmgb_ctx **out)
  {
    if (!out)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      return fail(nullptr, MFMGB_ERR_CUDA, "mfmgb_ctx_create: no CUDA device (%s)",
                  cudaGetErrorString(e));
    if (device < 0 || device >= count)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_ctx_create: device %d out of range [0,%d)",
                  device, count);
    mfmgb_ctx *ctx = new mfmgb_ctx();
    ctx->device = device;
    MFMGB_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    MFMGB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    if (prop.major < 10)
    {
      delete ctx;
      return fail(nullptr, MFMGB_ERR_CUDA,
                  "mfmgb_ctx_create: device is sm_%d%d; this library is built for sm_100a only",
                  prop.major, prop.minor);
    }
    if (stream)
    {
      ctx->stream = (cudaStream_t)stream;
      ctx->own_stream = false;
    }
    else
    {
      MFMGB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    ctx->red_capacity = 4096;
    MFMGB_CUDA(ctx, cudaMalloc(&ctx->red_partials, sizeof(double) * ctx->red_capacity * 4));
    MFMGB_CUDA(ctx, cudaMalloc(&ctx->red_result, sizeof(double) * 8));
    MFMGB_CUDA(ctx, cudaMallocHost(&ctx->red_result_host, sizeof(double) * 8));
    *out = ctx;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_ctx_destroy(mfmgb_ctx *ctx)
  {
    if (!ctx)
      return MFMGB_OK;
    cudaSetDevice(ctx->device);
    mfmgb_comm_finalize(ctx); // drops the communicator registered for this context, if any (no stale registry entry)
    cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t ev : ctx->prof_ev)
      cudaEventDestroy(ev);
    cudaFree(ctx->red_partials);
    cudaFree(ctx->red_result);
    cudaFreeHost(ctx->red_result_host);
    if (ctx->pinned)
      cudaFreeHost(ctx->pinned);
    if (ctx->own_stream)
      cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_ctx_synchronize(mfmgb_ctx *ctx)
  {
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return mfmgb_comm_check(ctx); // a kernel that gave up waiting for a peer GPU is reported here
  }

  MFMGB_API void *mfmgb_ctx_stream(mfmgb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

  MFMGB_API int64_t mfmgb_ctx_launch_count(mfmgb_ctx *ctx) { return ctx ? ctx->launches : 0; }

  // ctx may be NULL in the four mfmgb_dev_* calls (the reference's cuda_malloc & co. take no handle): errors then go to
  // the calling thread's message, copies and frees synchronise the whole device instead of the context's stream.
  MFMGB_API int mfmgb_dev_malloc(mfmgb_ctx *ctx, int64_t bytes, void **out)
  {
    MFMGB_REQUIRE(ctx, out && bytes >= 0, "mfmgb_dev_malloc: bad arguments");
    *out = nullptr;
    MFMGB_CUDA(ctx, cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 1)));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_dev_free(mfmgb_ctx *ctx, void *ptr)
  {
    if (ptr)
    {
      MFMGB_CUDA(ctx, ctx ? cudaStreamSynchronize(ctx->stream) : cudaDeviceSynchronize());
      MFMGB_CUDA(ctx, cudaFree(ptr));
    }
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_dev_upload(mfmgb_ctx *ctx, void *dst_dev, const void *src_host, int64_t bytes)
  {
    MFMGB_REQUIRE(ctx, bytes >= 0 && (bytes == 0 || (dst_dev && src_host)), "mfmgb_dev_upload: bad arguments");
    if (!ctx)
    {
      MFMGB_CUDA(ctx, cudaMemcpy(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice));
      return MFMGB_OK;
    }
    MFMGB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_dev_download(mfmgb_ctx *ctx, const void *src_dev, void *dst_host, int64_t bytes)
  {
    MFMGB_REQUIRE(ctx, bytes >= 0 && (bytes == 0 || (dst_host && src_dev)), "mfmgb_dev_download: bad arguments");
    if (!ctx)
    {
      MFMGB_CUDA(ctx, cudaMemcpy(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost));
      return MFMGB_OK;
    }
    MFMGB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_vec_alloc(mfmgb_ctx *ctx, int64_t n, double **out)
  {
    MFMGB_REQUIRE(ctx, ctx && out && n >= 0, "mfmgb_vec_alloc: bad arguments");
    *out = nullptr;
    // +2 doubles of slack so that 128-bit tail loads never leave the allocation
    MFMGB_CUDA(ctx, cudaMalloc(out, sizeof(double) * (size_t)(n + 2)));
    MFMGB_CUDA(ctx, cudaMemsetAsync(*out, 0, sizeof(double) * (size_t)(n + 2), ctx->stream));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_vec_free(mfmgb_ctx *ctx, double *v)
  {
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    if (v)
    {
      MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      MFMGB_CUDA(ctx, cudaFree(v));
    }
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_vec_upload(mfmgb_ctx *ctx, double *dst_dev, const double *src_host, int64_t n)
  {
    MFMGB_REQUIRE(ctx, ctx && (n == 0 || (dst_dev && src_host)), "mfmgb_vec_upload: bad arguments");
    MFMGB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, sizeof(double) * (size_t)n,
                                    cudaMemcpyHostToDevice, ctx->stream));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_vec_download(mfmgb_ctx *ctx, const double *src_dev, double *dst_host, int64_t n)
  {
    MFMGB_REQUIRE(ctx, ctx && (n == 0 || (dst_host && src_dev)), "mfmgb_vec_download: bad arguments");
    MFMGB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, sizeof(double) * (size_t)n,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_vec_fill(mfmgb_ctx *ctx, double *v, double value, int64_t n)
  {
    MFMGB_REQUIRE(ctx, ctx && (n == 0 || v), "mfmgb_vec_fill: bad arguments");
    return vec_fill(ctx, v, value, n);
  }

  MFMGB_API int mfmgb_vec_copy(mfmgb_ctx *ctx, double *dst, const double *src, int64_t n)
  {
    MFMGB_REQUIRE(ctx, ctx && (n == 0 || (dst && src)), "mfmgb_vec_copy: bad arguments");
    MFMGB_CUDA(ctx, cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_vec_axpy(mfmgb_ctx *ctx, double *y, double a, const double *x, int64_t n)
  {
    MFMGB_REQUIRE(ctx, ctx && (n == 0 || (y && x)), "mfmgb_vec_axpy: bad arguments");
    return vec_axpy(ctx, y, a, x, n);
  }

  MFMGB_API int mfmgb_vec_dot(mfmgb_ctx *ctx, const double *a, const double *b, int64_t n,
                              double *result_host)
  {
    MFMGB_REQUIRE(ctx, ctx && result_host && (n == 0 || (a && b)), "mfmgb_vec_dot: bad arguments");
    MFMGB_CHECK(vec_dot_async(ctx, a, b, n, ctx->red_result));
    MFMGB_CUDA(ctx, cudaMemcpyAsync(ctx->red_result_host, ctx->red_result, sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->stream));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *result_host = ctx->red_result_host[0];
    return MFMGB_OK;
  }
}

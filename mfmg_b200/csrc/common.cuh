// common.cuh -- context, error handling and small device helpers shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX 3: ranges are no-ops unless a tool (nsys, ncu) is attached

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "mfmg_b200.h"

namespace mfmgb
{
constexpr int kNumSMs = 148; // B200

inline std::string &tls_error()
{
  static thread_local std::string e;
  return e;
}
} // namespace mfmgb

struct mfmgb_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = mfmgb::kNumSMs;
  int64_t launches = 0;
  std::string error;
  // scratch for deterministic two-pass reductions
  double *red_partials = nullptr; // [red_capacity]
  double *red_result = nullptr;   // device [8]
  double *red_result_host = nullptr; // pinned [8]
  int red_capacity = 0;
  // pinned staging for *_host entry points
  double *pinned = nullptr;
  size_t pinned_bytes = 0;
  // objects holding instantiated CUDA graphs with NCCL nodes: they must be dropped before the communicator is
  // destroyed (ncclCommDestroy otherwise waits forever).  (owner, drop function)
  std::vector<std::pair<void *, void (*)(void *)>> graph_owners;
  // timeline marks (mfmgb_vcycle_timeline): CUDA events recorded on the compute stream between the pieces of a cycle;
  // inside a stream capture they become event-record nodes, so the times are those of the GRAPH replay
  bool prof_on = false;
  bool prof_capturing = false; // the marks are recorded inside a stream capture (event-record nodes)
  int prof_n = 0;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<std::string> prof_names;
};

namespace mfmgb
{
inline int fail(mfmgb_ctx *ctx, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->error = buf;
  tls_error() = buf;
  return code;
}

#define MFMGB_CUDA(ctx, call)                                                                       \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return mfmgb::fail((ctx), MFMGB_ERR_CUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,  \
                         cudaGetErrorString(e__));                                                  \
  } while (0)

#define MFMGB_CHECK(call)                                                                           \
  do                                                                                                \
  {                                                                                                 \
    int rc__ = (call);                                                                              \
    if (rc__ != MFMGB_OK)                                                                           \
      return rc__;                                                                                  \
  } while (0)

#define MFMGB_REQUIRE(ctx, cond, msg)                                                               \
  do                                                                                                \
  {                                                                                                 \
    if (!(cond))                                                                                    \
      return mfmgb::fail((ctx), MFMGB_ERR_INVALID, "%s:%d: %s", __FILE__, __LINE__, (msg));         \
  } while (0)

// launch bookkeeping: every kernel launch goes through this so gpu_launches is a real count
#define MFMGB_LAUNCHED(ctx)                                                                         \
  do                                                                                                \
  {                                                                                                 \
    (ctx)->launches++;                                                                              \
    MFMGB_CUDA((ctx), cudaGetLastError());                                                          \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// marks the END of the piece called `name` on the compute stream (no-op unless a timeline is being taken)
// NVTX range over a host entry point -- the counterpart of the reference's timer sections (hierarchy.hpp:36-47, 241, 263,
// 271: "Setup", "Apply", per-level sections), visible in nsys / ncu timelines
struct NvtxRange
{
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

inline void prof_mark(mfmgb_ctx *ctx, const char *name)
{
  if (!ctx->prof_on)
    return;
  if (ctx->prof_n >= (int)ctx->prof_ev.size())
  {
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess)
      return;
    ctx->prof_ev.push_back(ev);
    ctx->prof_names.emplace_back();
  }
  // inside a capture a plain cudaEventRecord only marks a dependency; the External flag makes it a real
  // event-record node whose timestamp is taken at every replay (the flag is invalid outside a capture)
  if (cudaEventRecordWithFlags(ctx->prof_ev[(size_t)ctx->prof_n], ctx->stream,
                               ctx->prof_capturing ? cudaEventRecordExternal : cudaEventRecordDefault) != cudaSuccess)
    return;
  ctx->prof_names[(size_t)ctx->prof_n] = name;
  ctx->prof_n++;
}

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ double ldg_f64(const double *p) { return __ldg(p); }

// streaming (read-once) loads: bypass L1 allocation so x stays cached
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_i32(const int *p)
{
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p)
{
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream_i32x4(const int *p)
{
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <int WIDTH>
__device__ __forceinline__ double subwarp_sum(double v)
{
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o, 32);
  return v;
}

// block-wide sum with a fixed tree (deterministic). BLOCK must be a multiple of 32, <= 1024.
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double *smem /* >= BLOCK/32 */)
{
  v = subwarp_sum<32>(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0)
    smem[w] = v;
  __syncthreads();
  double r = 0.;
  if (w == 0)
  {
    r = lane < BLOCK / 32 ? smem[lane] : 0.;
    r = subwarp_sum<32>(r);
  }
  return r; // valid in warp 0
}
} // namespace mfmgb

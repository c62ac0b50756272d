// comm.cuh -- NCCL communicator and halo plans for the row-partitioned levels (internal).
#pragma once
#include <nccl.h>

#include "common.cuh"

struct mfmgb_comm
{
  ncclComm_t nccl = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t stream = nullptr; // communication stream (halo exchange overlaps interior rows)
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};

struct mfmgb_halo
{
  int64_t n_owned = 0, n_ghost = 0;
  int n_neighbors = 0;
  std::vector<int> ranks;
  std::vector<int64_t> send_off, send_cnt, recv_off, recv_cnt; // offsets into sendbuf / the ghost tail
  std::vector<int64_t> send_first;                            // first local index when the list is contiguous
  bool contiguous = true;   // every send list is a contiguous index range => send straight from the vector
  int32_t *send_idx = nullptr; // device, concatenated
  double *sendbuf = nullptr;   // device
  int64_t n_send = 0;
};

namespace mfmgb
{
#define MFMGB_NCCL(ctx, call)                                                                       \
  do                                                                                                \
  {                                                                                                 \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess)                                                                         \
      return mfmgb::fail((ctx), MFMGB_ERR_NCCL, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,  \
                         ncclGetErrorString(r__));                                                  \
  } while (0)

mfmgb_comm *ctx_comm(mfmgb_ctx *ctx);
// start the exchange of v's ghost tail on the communication stream (after everything queued on the compute
// stream so far); halo_wait makes the compute stream wait for it.
int halo_start(mfmgb_ctx *ctx, const mfmgb_halo *h, double *v);
int halo_wait(mfmgb_ctx *ctx);
// in-stream (compute stream) sum over ranks of n doubles
int allreduce_sum(mfmgb_ctx *ctx, double *dev, int n);
// gather the rank-local slices [offsets[r], offsets[r+1]) of `full` so that every rank holds all of it
int allgather_slices(mfmgb_ctx *ctx, double *full, const std::vector<int64_t> &offsets);
} // namespace mfmgb

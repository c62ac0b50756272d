// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Replaces the reference's per-SpMV host MPI_Allgatherv of the WHOLE source vector
// (include/mfmg/cuda/sparse_matrix_device.templates.cuh:104-138, source/cuda/utils.cu:305-482) by a halo
// exchange of the boundary entries only: grouped ncclSend/ncclRecv on a communication stream that overlaps the
// interior rows of the SpMV; CG dot products use ncclAllReduce; the coarse right-hand side is all-gathered
// (the dense coarse solve is replicated).
#include <cstring>
#include <map>

#include "comm.cuh"

using namespace mfmgb;

namespace
{
std::map<mfmgb_ctx *, mfmgb_comm *> &registry()
{
  static std::map<mfmgb_ctx *, mfmgb_comm *> r;
  return r;
}

__global__ void __launch_bounds__(256)
    pack_kernel(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ v, double *__restrict__ buf)
{
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n)
    buf[i] = v[idx[i]];
}
} // namespace

namespace mfmgb
{
mfmgb_comm *ctx_comm(mfmgb_ctx *ctx)
{
  auto it = registry().find(ctx);
  return it == registry().end() ? nullptr : it->second;
}

int halo_start(mfmgb_ctx *ctx, const mfmgb_halo *h, double *v)
{
  mfmgb_comm *c = ctx_comm(ctx);
  if (!c)
    return fail(ctx, MFMGB_ERR_INVALID, "halo exchange requested but mfmgb_comm_init was not called");
  MFMGB_CUDA(ctx, cudaEventRecord(c->ev_ready, ctx->stream));
  MFMGB_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ev_ready, 0));
  if (!h->contiguous && h->n_send > 0)
  {
    pack_kernel<<<(unsigned)ceil_div(h->n_send, 256), 256, 0, c->stream>>>(h->n_send, h->send_idx, v, h->sendbuf);
    ctx->launches++;
    MFMGB_CUDA(ctx, cudaGetLastError());
  }
  MFMGB_NCCL(ctx, ncclGroupStart());
  for (int k = 0; k < h->n_neighbors; ++k)
  {
    if (h->send_cnt[k] > 0)
    {
      const double *src = h->contiguous ? v + h->send_first[k] : h->sendbuf + h->send_off[k];
      MFMGB_NCCL(ctx, ncclSend(src, (size_t)h->send_cnt[k], ncclDouble, h->ranks[k], c->nccl, c->stream));
    }
    if (h->recv_cnt[k] > 0)
      MFMGB_NCCL(ctx, ncclRecv(v + h->n_owned + h->recv_off[k], (size_t)h->recv_cnt[k], ncclDouble, h->ranks[k],
                               c->nccl, c->stream));
  }
  MFMGB_NCCL(ctx, ncclGroupEnd());
  MFMGB_CUDA(ctx, cudaEventRecord(c->ev_done, c->stream));
  return MFMGB_OK;
}

int halo_wait(mfmgb_ctx *ctx)
{
  mfmgb_comm *c = ctx_comm(ctx);
  MFMGB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_done, 0));
  return MFMGB_OK;
}

int allreduce_sum(mfmgb_ctx *ctx, double *dev, int n)
{
  mfmgb_comm *c = ctx_comm(ctx);
  if (!c || c->nranks == 1)
    return MFMGB_OK;
  MFMGB_NCCL(ctx, ncclAllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, c->nccl, ctx->stream));
  return MFMGB_OK;
}

int allgather_slices(mfmgb_ctx *ctx, double *full, const std::vector<int64_t> &off)
{
  mfmgb_comm *c = ctx_comm(ctx);
  if (!c || c->nranks == 1)
    return MFMGB_OK;
  const int nr = c->nranks, me = c->rank;
  bool equal = true;
  for (int r = 0; r < nr; ++r)
    equal = equal && (off[r + 1] - off[r] == off[1] - off[0]);
  if (equal)
  {
    MFMGB_NCCL(ctx, ncclAllGather(full + off[me], full, (size_t)(off[1] - off[0]), ncclDouble, c->nccl, ctx->stream));
    return MFMGB_OK;
  }
  MFMGB_NCCL(ctx, ncclGroupStart());
  for (int r = 0; r < nr; ++r)
  {
    if (r == me)
      continue;
    if (off[me + 1] > off[me])
      MFMGB_NCCL(ctx, ncclSend(full + off[me], (size_t)(off[me + 1] - off[me]), ncclDouble, r, c->nccl, ctx->stream));
    if (off[r + 1] > off[r])
      MFMGB_NCCL(ctx, ncclRecv(full + off[r], (size_t)(off[r + 1] - off[r]), ncclDouble, r, c->nccl, ctx->stream));
  }
  MFMGB_NCCL(ctx, ncclGroupEnd());
  return MFMGB_OK;
}
} // namespace mfmgb

extern "C"
{
  MFMGB_API int mfmgb_comm_unique_id(char *out128)
  {
    if (!out128)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_comm_unique_id: out is NULL");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess)
      return fail(nullptr, MFMGB_ERR_NCCL, "ncclGetUniqueId failed: %s", ncclGetErrorString(r));
    memcpy(out128, &id, 128);
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_comm_init(mfmgb_ctx *ctx, const char *id128, int nranks, int rank)
  {
    MFMGB_REQUIRE(ctx, ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "mfmgb_comm_init: bad arguments");
    MFMGB_REQUIRE(ctx, ctx_comm(ctx) == nullptr, "mfmgb_comm_init: already initialised");
    mfmgb_comm *c = new mfmgb_comm();
    c->nranks = nranks;
    c->rank = rank;
    MFMGB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    MFMGB_NCCL(ctx, ncclCommInitRank(&c->nccl, nranks, id, rank));
    MFMGB_CUDA(ctx, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    registry()[ctx] = c;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_comm_finalize(mfmgb_ctx *ctx)
  {
    mfmgb_comm *c = ctx_comm(ctx);
    if (!c)
      return MFMGB_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(c->stream);
    for (auto &o : ctx->graph_owners) // captured V-cycles reference the communicator
      o.second(o.first);
    ncclCommDestroy(c->nccl);
    cudaEventDestroy(c->ev_ready);
    cudaEventDestroy(c->ev_done);
    cudaStreamDestroy(c->stream);
    registry().erase(ctx);
    delete c;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_comm_rank(mfmgb_ctx *ctx)
  {
    mfmgb_comm *c = ctx_comm(ctx);
    return c ? c->rank : 0;
  }
  MFMGB_API int mfmgb_comm_size(mfmgb_ctx *ctx)
  {
    mfmgb_comm *c = ctx_comm(ctx);
    return c ? c->nranks : 1;
  }

  MFMGB_API int mfmgb_halo_create(mfmgb_ctx *ctx, int64_t n_owned, int64_t n_ghost, int n_neighbors,
                                  const int *neighbor_ranks, const int64_t *send_counts, const int32_t *send_indices,
                                  const int64_t *recv_counts, mfmgb_halo **out)
  {
    MFMGB_REQUIRE(ctx, ctx && out && n_owned >= 0 && n_ghost >= 0 && n_neighbors >= 0, "mfmgb_halo_create: bad arguments");
    MFMGB_REQUIRE(ctx, n_neighbors == 0 || (neighbor_ranks && send_counts && recv_counts),
                  "mfmgb_halo_create: NULL neighbour arrays");
    mfmgb_halo *h = new mfmgb_halo();
    h->n_owned = n_owned;
    h->n_ghost = n_ghost;
    h->n_neighbors = n_neighbors;
    int64_t so = 0, ro = 0;
    for (int k = 0; k < n_neighbors; ++k)
    {
      h->ranks.push_back(neighbor_ranks[k]);
      h->send_off.push_back(so);
      h->send_cnt.push_back(send_counts[k]);
      h->recv_off.push_back(ro);
      h->recv_cnt.push_back(recv_counts[k]);
      int64_t first = send_counts[k] > 0 ? send_indices[so] : 0;
      for (int64_t i = 0; i < send_counts[k]; ++i)
      {
        const int64_t idx = send_indices[so + i];
        if (idx < 0 || idx >= n_owned)
        {
          delete h;
          return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_halo_create: send index %lld out of range", (long long)idx);
        }
        if (idx != first + i)
          h->contiguous = false;
      }
      h->send_first.push_back(first);
      so += send_counts[k];
      ro += recv_counts[k];
    }
    if (ro != n_ghost)
    {
      delete h;
      return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_halo_create: recv counts sum to %lld, expected n_ghost = %lld",
                  (long long)ro, (long long)n_ghost);
    }
    h->n_send = so;
    if (so > 0)
    {
      MFMGB_CUDA(ctx, cudaMalloc(&h->send_idx, sizeof(int32_t) * (size_t)so));
      MFMGB_CUDA(ctx, cudaMemcpy(h->send_idx, send_indices, sizeof(int32_t) * (size_t)so, cudaMemcpyHostToDevice));
      MFMGB_CUDA(ctx, cudaMalloc(&h->sendbuf, sizeof(double) * (size_t)so));
    }
    *out = h;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_halo_destroy(mfmgb_ctx *ctx, mfmgb_halo *h)
  {
    if (!h)
      return MFMGB_OK;
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(h->send_idx);
    cudaFree(h->sendbuf);
    delete h;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_halo_exchange(mfmgb_ctx *ctx, const mfmgb_halo *h, double *v)
  {
    MFMGB_REQUIRE(ctx, ctx && h && v, "mfmgb_halo_exchange: bad arguments");
    MFMGB_CHECK(halo_start(ctx, h, v));
    return halo_wait(ctx);
  }

  MFMGB_API int mfmgb_allreduce_sum(mfmgb_ctx *ctx, double *dev, int n)
  {
    MFMGB_REQUIRE(ctx, ctx && dev && n >= 0, "mfmgb_allreduce_sum: bad arguments");
    return allreduce_sum(ctx, dev, n);
  }
}

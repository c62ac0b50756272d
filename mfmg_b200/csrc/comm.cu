// comm.cu -- multi-GPU plumbing: one process per GPU; ghost entries and small reductions travel over NVLink / NVSwitch
// PEER MEMORY written by our own kernels, NCCL is the bootstrap (and the fallback when CUDA IPC is unavailable).
//
// Replaces the reference's per-SpMV host MPI_Allgatherv of the WHOLE source vector
// (include/mfmg/cuda/sparse_matrix_device.templates.cuh:104-138, source/cuda/utils.cu:305-482) by a halo exchange of
// the boundary entries only, overlapped with the interior rows of the SpMV:
//   * peer path (default): every rank exports one window of device memory with CUDA IPC; the push kernel gathers the
//     boundary entries of v and STORES them into the neighbour's mailbox through the mapped pointer, fences, and raises
//     a sequence flag there (st.release.sys); the neighbour's wait kernel spins on its own flag (ld.acquire.sys), then
//     copies the mailbox into the ghost tail.  Mailboxes are double-buffered by the parity of the exchange count:
//     a sender can only be one exchange ahead of its neighbour (its next-but-one push needs the neighbour's next
//     push, which the neighbour issues after it has emptied the mailbox), so no acknowledgement travels back.
//     Sequence numbers live in device memory and are advanced by the kernels: CUDA-graph replay == eager launches.
//   * small reductions (CG dots, the separator right-hand side of the coarse solve): ONE kernel -- store the local
//     values into every rank's slot, flag, wait for all flags, sum the slots in rank order (identical bits on all ranks).
//   * NCCL path (MFMGB_PEER=0, or IPC mapping failed on some rank): grouped ncclSend/ncclRecv and ncclAllReduce.
// No kernel of a rank ever waits for something a LATER kernel of the same rank produces, and pushes never wait:
// every wait is satisfied by work the peer issues unconditionally (no deadlock across GPUs).
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "comm.cuh"
#include "peer.cuh"

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

struct PushArgs
{
  const mfmgb_halo::Link *links;
  int n_links, nranks, rank;
  const int32_t *send_idx; // concatenated local indices (NULL: contiguous ranges starting at send_first)
  const double *v;
  unsigned char *const *base; // mapped windows of all ranks
  size_t box_off, flag_off;
  long long box_cap;
  unsigned long long *seq; // [0] pushes so far
  unsigned int *done;      // [0] CTA completion counter
};

// boundary entries of v -> the neighbours' mailboxes (remote stores over NVLink), then one flag per neighbour
__global__ void __launch_bounds__(256) halo_push_kernel(const PushArgs a)
{
  const unsigned long long s = a.seq[0] + 1; // (stable: only the last CTA to finish advances it)
  const size_t slot = ((size_t)(s & 1ull) * (size_t)a.nranks + (size_t)a.rank) * (size_t)a.box_cap;
  for (int k = 0; k < a.n_links; ++k)
  {
    const mfmgb_halo::Link l = a.links[k];
    double *dst = reinterpret_cast<double *>(a.base[l.rank] + a.box_off) + slot;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < l.send_cnt; i += (long long)gridDim.x * 256)
      dst[i] = a.send_idx ? a.v[a.send_idx[l.send_off + i]] : a.v[l.send_first + i];
  }
  // one system fence per CTA (thread 0, after the barrier that orders the CTA's remote stores before it), not one per
  // thread: every fence waits for the NVLink write acknowledgements of its SM
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence_system();
    const unsigned int prev = atomicAdd(&a.done[0], 1u);
    if (prev == gridDim.x - 1)
    {
      a.done[0] = 0;
      __threadfence_system();
      for (int k = 0; k < a.n_links; ++k)
      {
        unsigned long long *flag = reinterpret_cast<unsigned long long *>(a.base[a.links[k].rank] + a.flag_off) +
                                   ((s & 1ull) * (unsigned long long)a.nranks + (unsigned long long)a.rank);
        st_release_sys(flag, s);
      }
      a.seq[0] = s;
    }
  }
}

struct WaitArgs
{
  const mfmgb_halo::Link *links;
  int n_links, nranks;
  double *v;
  long long n_owned;
  unsigned char *local; // this rank's window
  size_t box_off, flag_off;
  long long box_cap;
  unsigned long long *seq; // [1] waits so far
  unsigned int *done;      // [1]
  unsigned long long timeout_ns;
  int *err;
};

// wait for every neighbour's flag of this exchange, then mailbox -> ghost tail of v
__global__ void __launch_bounds__(256) halo_wait_kernel(const WaitArgs a)
{
  const unsigned long long s = a.seq[1] + 1;
  const size_t par = (size_t)(s & 1ull);
  if ((int)threadIdx.x < a.n_links)
  {
    const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(a.local + a.flag_off) +
                                     (par * (size_t)a.nranks + (size_t)a.links[threadIdx.x].rank);
    wait_flag(flag, s, a.timeout_ns, a.err);
  }
  __syncthreads();
  for (int k = 0; k < a.n_links; ++k)
  {
    const mfmgb_halo::Link l = a.links[k];
    const double *src = reinterpret_cast<const double *>(a.local + a.box_off) +
                        (par * (size_t)a.nranks + (size_t)l.rank) * (size_t)a.box_cap;
    double *dst = a.v + a.n_owned + l.recv_off;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < l.recv_cnt; i += (long long)gridDim.x * 256)
      dst[i] = __ldcg(src + i);
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    const unsigned int prev = atomicAdd(&a.done[1], 1u);
    if (prev == gridDim.x - 1)
    {
      a.done[1] = 0;
      a.seq[1] = s;
    }
  }
}

// one-kernel all-reduce (sum) of n <= cap doubles over all ranks through peer memory; one CTA
__global__ void __launch_bounds__(1024) peer_allreduce_kernel(double *__restrict__ buf, int n, const PeerAllreduceArgs a)
{
  peer_allreduce_cta(buf, n, a);
}

int env_int(const char *name, int dflt)
{
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// collective max over the ranks of one 64-bit value (setup time; synchronises)
int agree_max(mfmgb_ctx *ctx, mfmgb_comm *c, long long *value)
{
  long long *dev = nullptr;
  MFMGB_CUDA(ctx, cudaMalloc(&dev, sizeof(long long)));
  MFMGB_CUDA(ctx, cudaMemcpy(dev, value, sizeof(long long), cudaMemcpyHostToDevice));
  MFMGB_NCCL(ctx, ncclAllReduce(dev, dev, 1, ncclInt64, ncclMax, c->nccl, c->stream));
  MFMGB_CUDA(ctx, cudaStreamSynchronize(c->stream));
  MFMGB_CUDA(ctx, cudaMemcpy(value, dev, sizeof(long long), cudaMemcpyDeviceToHost));
  cudaFree(dev);
  return MFMGB_OK;
}

// Export this rank's window, map everybody else's.  Any failure on any rank switches the whole communicator to NCCL.
int peer_setup(mfmgb_ctx *ctx, mfmgb_comm *c)
{
  mfmgb_peer &p = c->peer;
  p.enabled = false;
  if (c->nranks < 2 || env_int("MFMGB_PEER", 1) == 0)
    return MFMGB_OK;
  p.bytes = (size_t)std::max(8, env_int("MFMGB_PEER_WINDOW_MB", 96)) << 20;
  p.timeout_ns = (unsigned long long)std::max(1, env_int("MFMGB_PEER_TIMEOUT_MS", 20000)) * 1000000ull;
  long long ok = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&p.local, p.bytes) != cudaSuccess || cudaMemset(p.local, 0, p.bytes) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine, p.local) != cudaSuccess)
    ok = 0;
  cudaGetLastError();
  cudaDeviceSynchronize();
  // all-gather the 64-byte handles with NCCL (device staging)
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
  char *dsend = nullptr, *dall = nullptr;
  MFMGB_CUDA(ctx, cudaMalloc(&dsend, 64));
  MFMGB_CUDA(ctx, cudaMalloc(&dall, 64 * (size_t)c->nranks));
  MFMGB_CUDA(ctx, cudaMemcpy(dsend, &mine, 64, cudaMemcpyHostToDevice));
  MFMGB_NCCL(ctx, ncclAllGather(dsend, dall, 64, ncclChar, c->nccl, c->stream));
  MFMGB_CUDA(ctx, cudaStreamSynchronize(c->stream));
  MFMGB_CUDA(ctx, cudaMemcpy(all.data(), dall, 64 * (size_t)c->nranks, cudaMemcpyDeviceToHost));
  cudaFree(dsend);
  cudaFree(dall);
  long long all_ok = -ok; // max of -ok == -(min of ok)
  MFMGB_CHECK(agree_max(ctx, c, &all_ok));
  p.base.assign((size_t)c->nranks, nullptr);
  if (all_ok == -1)
  {
    for (int r = 0; r < c->nranks && ok; ++r)
    {
      if (r == c->rank)
      {
        p.base[(size_t)r] = p.local;
        continue;
      }
      void *mapped = nullptr;
      if (cudaIpcOpenMemHandle(&mapped, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      {
        cudaGetLastError();
        ok = 0;
      }
      p.base[(size_t)r] = static_cast<unsigned char *>(mapped);
    }
    all_ok = -ok;
    MFMGB_CHECK(agree_max(ctx, c, &all_ok));
  }
  if (all_ok != -1)
  {
    for (int r = 0; r < c->nranks; ++r)
      if (r != c->rank && p.base[(size_t)r])
        cudaIpcCloseMemHandle(p.base[(size_t)r]);
    cudaFree(p.local);
    p.local = nullptr;
    p.base.clear();
    return MFMGB_OK; // NCCL transport
  }
  MFMGB_CUDA(ctx, cudaMalloc(&p.base_dev, sizeof(unsigned char *) * (size_t)c->nranks));
  MFMGB_CUDA(ctx, cudaMemcpy(p.base_dev, p.base.data(), sizeof(unsigned char *) * (size_t)c->nranks,
                             cudaMemcpyHostToDevice));
  MFMGB_CUDA(ctx, cudaHostAlloc(&p.err_host, sizeof(int), cudaHostAllocMapped));
  *p.err_host = 0;
  MFMGB_CUDA(ctx, cudaHostGetDevicePointer(&p.err_dev, p.err_host, 0));
  MFMGB_CUDA(ctx, cudaMalloc(&p.ar_seq, sizeof(unsigned long long)));
  MFMGB_CUDA(ctx, cudaMemset(p.ar_seq, 0, sizeof(unsigned long long)));
  p.enabled = true;
  // the small all-reduce channel
  p.ar_cap = std::max(64, env_int("MFMGB_PEER_ALLREDUCE_CAP", 8192));
  size_t off = 0;
  MFMGB_CHECK(peer_alloc(ctx, (size_t)2 * (size_t)c->nranks * (size_t)p.ar_cap * sizeof(double), &off));
  p.ar_off = off;
  MFMGB_CHECK(peer_alloc(ctx, (size_t)2 * (size_t)c->nranks * sizeof(unsigned long long), &off));
  p.ar_flag_off = off;
  MFMGB_CUDA(ctx, cudaDeviceSynchronize());
  return MFMGB_OK;
}

void peer_teardown(mfmgb_comm *c)
{
  mfmgb_peer &p = c->peer;
  if (!p.local)
    return;
  for (int r = 0; r < c->nranks; ++r)
    if (r != c->rank && r < (int)p.base.size() && p.base[(size_t)r])
      cudaIpcCloseMemHandle(p.base[(size_t)r]);
  cudaFree(p.local);
  cudaFree(p.base_dev);
  cudaFree(p.ar_seq);
  if (p.err_host)
    cudaFreeHost(p.err_host);
  p = mfmgb_peer();
}
} // namespace

namespace mfmgb
{
int comm_agree_max(mfmgb_ctx *ctx, long long *value) { return agree_max(ctx, ctx_comm(ctx), value); }

mfmgb_comm *ctx_comm(mfmgb_ctx *ctx)
{
  auto it = registry().find(ctx);
  return it == registry().end() ? nullptr : it->second;
}

int peer_alloc(mfmgb_ctx *ctx, size_t bytes, size_t *offset)
{
  *offset = (size_t)-1;
  mfmgb_comm *c = ctx_comm(ctx);
  if (!c || !c->peer.local)
    return MFMGB_OK;
  long long agreed = (long long)((bytes + 255) & ~(size_t)255);
  MFMGB_CHECK(agree_max(ctx, c, &agreed));
  if (c->peer.used + (size_t)agreed > c->peer.bytes)
    return MFMGB_OK; // full: the caller falls back to NCCL (every rank takes the same decision)
  *offset = c->peer.used;
  c->peer.used += (size_t)agreed;
  return MFMGB_OK;
}

int peer_error(mfmgb_ctx *ctx)
{
  mfmgb_comm *c = ctx_comm(ctx);
  return c && c->peer.err_host ? *c->peer.err_host : 0;
}

static PushArgs make_push_args(const mfmgb_comm *c, const mfmgb_halo *h, const double *v)
{
  PushArgs a;
  a.links = h->links;
  a.n_links = h->n_neighbors;
  a.nranks = c->nranks;
  a.rank = c->rank;
  a.send_idx = h->contiguous ? nullptr : h->send_idx;
  a.v = v;
  a.base = c->peer.base_dev;
  a.box_off = h->box_off;
  a.flag_off = h->flag_off;
  a.box_cap = (long long)h->box_cap;
  a.seq = h->seq;
  a.done = h->done;
  return a;
}
static unsigned push_grid(const mfmgb_halo *h)
{
  int64_t widest = 1;
  for (int k = 0; k < h->n_neighbors; ++k)
    widest = std::max(widest, h->send_cnt[k]);
  return (unsigned)std::min<int64_t>(32, ceil_div(widest, 256 * 4));
}

int halo_start(mfmgb_ctx *ctx, const mfmgb_halo *h, double *v)
{
  mfmgb_comm *c = ctx_comm(ctx);
  if (!c)
    return fail(ctx, MFMGB_ERR_INVALID, "halo exchange requested but mfmgb_comm_init was not called");
  MFMGB_CUDA(ctx, cudaEventRecord(c->ev_ready, ctx->stream));
  MFMGB_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ev_ready, 0));
  if (h->peer)
  {
    // our own kernel stores the boundary entries into the neighbours' mailboxes over NVLink and raises their flags
    halo_push_kernel<<<push_grid(h), 256, 0, c->stream>>>(make_push_args(c, h, v));
    ctx->launches++;
    MFMGB_CUDA(ctx, cudaGetLastError());
    MFMGB_CUDA(ctx, cudaEventRecord(c->ev_done, c->stream));
    return MFMGB_OK;
  }
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

bool halo_can_fuse(mfmgb_ctx *ctx, const mfmgb_halo *h)
{
  static const bool enabled = [] {
    const char *v = getenv("MFMGB_HALO_FUSED");
    return !(v && v[0] == '0');
  }();
  mfmgb_comm *c = ctx_comm(ctx);
  return enabled && c && h->peer && c->peer.enabled;
}

int halo_push_inline(mfmgb_ctx *ctx, const mfmgb_halo *h, const double *v)
{
  mfmgb_comm *c = ctx_comm(ctx);
  halo_push_kernel<<<push_grid(h), 256, 0, ctx->stream>>>(make_push_args(c, h, v));
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

// process-wide launch parameters of the exchange kernels (mfmgb_tunable_set: measurement aid; defaults from the
// environment variables MFMGB_<NAME IN CAPITALS>)
struct Tunable
{
  const char *name;
  long long value;
  bool init;
};
static Tunable g_tunables[] = {
    {"halo_push_ctas", 16, false},   // CTAs of the fused A-kernel that store the boundary plane(s) into the neighbours
    {"halo_push_penalty", 0, false}, // interior tiles such a CTA is spared (measured: no effect, profiles/r02_probe_weak_n2_push_params.txt)
};

static Tunable *find_tunable(const char *name)
{
  for (Tunable &t : g_tunables)
    if (name && strcmp(t.name, name) == 0)
    {
      if (!t.init)
      {
        std::string env = std::string("MFMGB_") + t.name;
        for (char &ch : env)
          ch = (char)toupper((unsigned char)ch);
        const char *v = getenv(env.c_str());
        if (v && *v)
          t.value = atoll(v);
        t.init = true;
      }
      return &t;
    }
  return nullptr;
}

long long tunable(const char *name)
{
  Tunable *t = find_tunable(name);
  return t ? t->value : 0;
}

void halo_ghost_args(mfmgb_ctx *ctx, const mfmgb_halo *h, int64_t blo, int64_t bhi, GhostArgs *g)
{
  mfmgb_comm *c = ctx_comm(ctx);
  g->enabled = 1;
  g->n_links = h->n_neighbors;
  g->nranks = c->nranks;
  g->n_owned = (long long)h->n_owned;
  g->blo = (long long)blo;
  g->bhi = (long long)bhi;
  g->local = c->peer.local;
  g->box_off = h->box_off;
  g->flag_off = h->flag_off;
  g->box_cap = (long long)h->box_cap;
  g->links = h->ghost_links;
  g->seq = h->seq;
  g->done = h->done;
  g->timeout_ns = c->peer.timeout_ns;
  g->err = c->peer.err_dev;
  // push half inside the consumer kernel (MFMGB_HALO_PUSH_IN_KERNEL=0: separate push kernel on the compute stream)
  static const bool push_in_kernel = [] {
    const char *v = getenv("MFMGB_HALO_PUSH_IN_KERNEL");
    return !(v && v[0] == '0');
  }();
  g->n_push_ctas = push_in_kernel ? (int)tunable("halo_push_ctas") : 0;
  g->push_penalty = (int)tunable("halo_push_penalty");
  g->rank = c->rank;
  g->base = c->peer.base_dev;
  g->send_idx = h->contiguous ? nullptr : h->send_idx;
}

int halo_wait(mfmgb_ctx *ctx, const mfmgb_halo *h, double *v)
{
  mfmgb_comm *c = ctx_comm(ctx);
  // (peer path too: the push kernel reads v; later kernels of the compute stream may overwrite it)
  MFMGB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_done, 0));
  if (!h->peer)
    return MFMGB_OK;
  WaitArgs a;
  a.links = h->links;
  a.n_links = h->n_neighbors;
  a.nranks = c->nranks;
  a.v = v;
  a.n_owned = (long long)h->n_owned;
  a.local = c->peer.local;
  a.box_off = h->box_off;
  a.flag_off = h->flag_off;
  a.box_cap = (long long)h->box_cap;
  a.seq = h->seq;
  a.done = h->done;
  a.timeout_ns = c->peer.timeout_ns;
  a.err = c->peer.err_dev;
  int64_t widest = 1;
  for (int k = 0; k < h->n_neighbors; ++k)
    widest = std::max(widest, h->recv_cnt[k]);
  const unsigned grid = (unsigned)std::min<int64_t>(32, ceil_div(widest, 256 * 4));
  halo_wait_kernel<<<grid, 256, 0, ctx->stream>>>(a);
  ctx->launches++;
  MFMGB_CUDA(ctx, cudaGetLastError());
  return MFMGB_OK;
}

int allreduce_sum(mfmgb_ctx *ctx, double *dev, int n)
{
  mfmgb_comm *c = ctx_comm(ctx);
  if (!c || c->nranks == 1 || n <= 0)
    return MFMGB_OK;
  const mfmgb_peer &p = c->peer;
  if (p.enabled && n <= p.ar_cap)
  {
    const int threads = n >= 1024 ? 1024 : std::max(32, ((std::max(n, c->nranks) + 31) / 32) * 32);
    peer_allreduce_kernel<<<1, threads, 0, ctx->stream>>>(dev, n, peer_allreduce_args(c));
    ctx->launches++;
    MFMGB_CUDA(ctx, cudaGetLastError());
    return MFMGB_OK;
  }
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
    // the communication stream has the highest priority: its few CTAs (push kernels) are scheduled ahead of the
    // persistent CTAs of the interior-row kernel launched at the same time on the compute stream
    int prio_least = 0, prio_greatest = 0;
    MFMGB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    MFMGB_CUDA(ctx, cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
    MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    registry()[ctx] = c;
    if (nranks > 1)
      MFMGB_CHECK(peer_setup(ctx, c)); // collective; leaves c->peer.enabled == false when CUDA IPC is unavailable
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
    peer_teardown(c);
    ncclCommDestroy(c->nccl);
    cudaEventDestroy(c->ev_ready);
    cudaEventDestroy(c->ev_done);
    cudaStreamDestroy(c->stream);
    registry().erase(ctx);
    delete c;
    return MFMGB_OK;
  }

  MFMGB_API const char *mfmgb_comm_transport(mfmgb_ctx *ctx)
  {
    mfmgb_comm *c = ctx_comm(ctx);
    if (!c || c->nranks < 2)
      return "none (single GPU)";
    return c->peer.enabled ? "NVLink peer-memory stores by our own kernels (CUDA IPC window, flag-synchronised)"
                           : "NCCL send/recv + all-reduce";
  }

  MFMGB_API int mfmgb_comm_check(mfmgb_ctx *ctx)
  {
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    if (peer_error(ctx))
      return fail(ctx, MFMGB_ERR_NCCL, "a kernel gave up waiting for a peer GPU (MFMGB_PEER_TIMEOUT_MS): the ranks do "
                                       "not execute the same sequence of exchanges, or a peer died");
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
    // peer-memory mailboxes (collective when the communicator has a mapped window: every rank creates its plans in
    // the same order)
    mfmgb_comm *c = ctx_comm(ctx);
    if (c && c->peer.enabled)
    {
      int64_t cap = 1;
      for (int k = 0; k < n_neighbors; ++k)
        cap = std::max(cap, std::max(send_counts[k], recv_counts[k]));
      long long agreed_cap = (long long)((cap + 31) & ~(int64_t)31);
      MFMGB_CHECK(agree_max(ctx, c, &agreed_cap));
      size_t box = 0, flag = 0;
      MFMGB_CHECK(peer_alloc(ctx, (size_t)2 * (size_t)c->nranks * (size_t)agreed_cap * sizeof(double), &box));
      if (box != (size_t)-1)
        MFMGB_CHECK(peer_alloc(ctx, (size_t)2 * (size_t)c->nranks * sizeof(unsigned long long), &flag));
      if (box != (size_t)-1 && flag != (size_t)-1)
      {
        h->peer = true;
        h->box_off = box;
        h->flag_off = flag;
        h->box_cap = agreed_cap;
        std::vector<mfmgb_halo::Link> links((size_t)std::max(n_neighbors, 1));
        for (int k = 0; k < n_neighbors; ++k)
          links[(size_t)k] = {h->ranks[(size_t)k], (long long)h->send_off[(size_t)k], (long long)h->send_cnt[(size_t)k],
                              (long long)h->send_first[(size_t)k], (long long)h->recv_off[(size_t)k],
                              (long long)h->recv_cnt[(size_t)k]};
        MFMGB_CUDA(ctx, cudaMalloc(&h->links, sizeof(mfmgb_halo::Link) * links.size()));
        MFMGB_CUDA(ctx, cudaMemcpy(h->links, links.data(), sizeof(mfmgb_halo::Link) * links.size(), cudaMemcpyHostToDevice));
        std::vector<GhostLink> gl(links.size());
        for (int k = 0; k < n_neighbors; ++k)
          gl[(size_t)k] = {h->ranks[(size_t)k],          (long long)h->recv_off[(size_t)k],
                           (long long)h->recv_cnt[(size_t)k], (long long)h->send_off[(size_t)k],
                           (long long)h->send_cnt[(size_t)k], (long long)h->send_first[(size_t)k]};
        MFMGB_CUDA(ctx, cudaMalloc(&h->ghost_links, sizeof(GhostLink) * gl.size()));
        MFMGB_CUDA(ctx, cudaMemcpy(h->ghost_links, gl.data(), sizeof(GhostLink) * gl.size(), cudaMemcpyHostToDevice));
        MFMGB_CUDA(ctx, cudaMalloc(&h->seq, sizeof(unsigned long long) * 2));
        MFMGB_CUDA(ctx, cudaMemset(h->seq, 0, sizeof(unsigned long long) * 2));
        MFMGB_CUDA(ctx, cudaMalloc(&h->done, sizeof(unsigned int) * 2));
        MFMGB_CUDA(ctx, cudaMemset(h->done, 0, sizeof(unsigned int) * 2));
      }
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
    cudaFree(h->links);
    cudaFree(h->ghost_links);
    cudaFree(h->seq);
    cudaFree(h->done);
    delete h;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_halo_exchange(mfmgb_ctx *ctx, const mfmgb_halo *h, double *v)
  {
    MFMGB_REQUIRE(ctx, ctx && h && v, "mfmgb_halo_exchange: bad arguments");
    MFMGB_CHECK(halo_start(ctx, h, v));
    return halo_wait(ctx, h, v);
  }

  MFMGB_API int mfmgb_allreduce_sum(mfmgb_ctx *ctx, double *dev, int n)
  {
    MFMGB_REQUIRE(ctx, ctx && dev && n >= 0, "mfmgb_allreduce_sum: bad arguments");
    return allreduce_sum(ctx, dev, n);
  }
}

extern "C"
{
  MFMGB_API int mfmgb_tunable_set(const char *name, long long value)
  {
    mfmgb::Tunable *t = mfmgb::find_tunable(name);
    if (!t)
      return MFMGB_ERR_INVALID;
    t->value = value;
    return MFMGB_OK;
  }

  MFMGB_API long long mfmgb_tunable_get(const char *name)
  {
    mfmgb::Tunable *t = mfmgb::find_tunable(name);
    return t ? t->value : -1;
  }
}

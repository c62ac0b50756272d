// comm.cuh -- communicator, peer-memory window and halo plans for the row-partitioned levels (internal).
#pragma once
#include <nccl.h>

#include "common.cuh"
#include "csr.cuh"

// Peer-memory window (NVLink / NVSwitch): one cudaMalloc'ed block per rank, exported with CUDA IPC and mapped by every
// other rank of the node at mfmgb_comm_init.  Sub-allocations are SYMMETRIC -- every rank performs the same sequence
// of peer_alloc calls with the same sizes -- so "offset o in rank q's window" is base[q] + o on every rank.
struct mfmgb_peer
{
  bool enabled = false;
  size_t bytes = 0, used = 0;
  unsigned char *local = nullptr;       // this rank's window
  std::vector<unsigned char *> base;    // [nranks] mapped base of every rank's window (base[rank] == local)
  // small all-reduce channel (PCG dots, the separator right-hand side of the coarse solve): per parity, one slot per
  // source rank
  size_t ar_off = 0;                    // window offset of [2][nranks][ar_cap] doubles
  size_t ar_flag_off = 0;               // window offset of [2][nranks] uint64 flags
  int ar_cap = 0;                       // doubles per slot
  unsigned long long *ar_seq = nullptr; // device: all-reduces issued so far (the kernel increments it)
  unsigned char **base_dev = nullptr;   // device copy of base[]
  int *err_host = nullptr;              // mapped pinned: set by a kernel whose wait for a peer timed out
  int *err_dev = nullptr;               // device alias of err_host
  unsigned long long timeout_ns = 0;
};

struct mfmgb_comm
{
  ncclComm_t nccl = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t stream = nullptr; // communication stream (halo exchange overlaps interior rows), highest priority
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  mfmgb_peer peer;
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
  // peer-memory form (comm.cu): the sender stores its boundary entries straight into the receiver's mailbox over
  // NVLink and raises a flag; the receiver's wait kernel copies the mailbox into the ghost tail.  Mailboxes are
  // double-buffered by the parity of the exchange count and indexed by SOURCE RANK, symmetric across the ranks.
  bool peer = false;
  size_t box_off = 0, flag_off = 0; // window offsets: [2][nranks][box_cap] doubles, [2][nranks] uint64
  int64_t box_cap = 0;
  unsigned long long *seq = nullptr; // device [2]: exchanges pushed / exchanges waited for so far
  struct Link                        // device array [n_neighbors]
  {
    int rank;
    long long send_off, send_cnt, send_first, recv_off, recv_cnt;
  };
  Link *links = nullptr;
  mfmgb::GhostLink *ghost_links = nullptr;  // device [n_neighbors]: the same links as the fused consumer kernel reads them
  unsigned int *done = nullptr;      // device [2]: CTA completion counters of the push / wait kernels
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
int halo_wait(mfmgb_ctx *ctx, const mfmgb_halo *h, double *v);
// fused form for levels served by the tile kernel over peer memory: the push kernel runs on the COMPUTE stream right
// after the producer of v (no fork / join), and the consumer kernel waits for the flags itself (GhostArgs, csr.cuh)
bool halo_can_fuse(mfmgb_ctx *ctx, const mfmgb_halo *h);
int halo_push_inline(mfmgb_ctx *ctx, const mfmgb_halo *h, const double *v);
void halo_ghost_args(mfmgb_ctx *ctx, const mfmgb_halo *h, int64_t blo, int64_t bhi, GhostArgs *g);
// in-stream (compute stream) sum over ranks of n doubles: one kernel over peer memory when the window is mapped and
// n fits a slot (every rank sums the contributions in rank order: identical bits on all ranks), else ncclAllReduce
int allreduce_sum(mfmgb_ctx *ctx, double *dev, int n);
// gather the rank-local slices [offsets[r], offsets[r+1]) of `full` so that every rank holds all of it
int allgather_slices(mfmgb_ctx *ctx, double *full, const std::vector<int64_t> &offsets);
// symmetric sub-allocation of the peer window (collective: every rank calls it in the same order; the size is the
// maximum over the ranks).  Returns MFMGB_OK and *offset, or sets *offset = (size_t)-1 when the window is full / off.
int peer_alloc(mfmgb_ctx *ctx, size_t bytes, size_t *offset);
// collective max over the ranks of one 64-bit value (setup time; synchronises)
int comm_agree_max(mfmgb_ctx *ctx, long long *value);
// non-zero when a kernel of this context gave up waiting for a peer (checked at synchronisation points)
int peer_error(mfmgb_ctx *ctx);
} // namespace mfmgb

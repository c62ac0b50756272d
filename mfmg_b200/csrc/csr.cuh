// csr.cuh -- device CSR matrix and the SpMV family with fused epilogues (internal C++ API).
#pragma once
#include "common.cuh"

struct mfmgb_csr
{
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  double *val = nullptr;  // [nnz + pad]
  int32_t *col = nullptr; // [nnz + pad]
  void *rowptr = nullptr; // int32[n_rows+1] or int64[n_rows+1]
  bool off64 = false;
  bool owns = true;
  int lanes = 8;          // lanes per row used by the vector-CSR kernels
  int lanes_override = 0; // 0 = automatic
  int device = 0;
  // tile-streamed kernel (csr_tile.cu)
  bool padded = false;        // arrays carry slack for 16-byte-granular bulk copies (uploads do; adopted arrays do not)
  bool aligned16 = true;      // all three arrays start on 16-byte boundaries (bulk copies need it)
  int64_t tile_cap[6] = {0, 0, 0, 0, 0, 0}; // widest aligned nnz span of a tile, per lanes = 1, 2, 4, 8, 16, 32
  // rows [0, tile_rows[slot]) are served by the tile kernel: all rows when the arrays carry slack; for adopted arrays
  // the tiles whose 16-byte-granular copies would leave the allocations (the last one or two) go to the direct-load
  // kernel, which sums in the same order (bit-identical)
  int64_t tile_rows[6] = {0, 0, 0, 0, 0, 0};
  bool tile_ok = false;       // the tile kernel can serve this matrix with the current lanes
  int tile_stages = 0, tile_ctas = 0;
  int kernel_override = -1;   // -1 = automatic, 0 = vector-CSR (csr.cu), 1 = tile-streamed (csr_tile.cu)
};

namespace mfmgb
{
enum class Epi : int
{
  Spmv = 0,   // y = A x
  Resid = 1,  // y = A x - b                       (hierarchy.hpp:284-286)
  Jacobi = 2, // y = xin - omega * dinv * (A x - b) (cuda_smoother.cu:49-59)
  Sub = 3     // y = y - A x                       (hierarchy.hpp:297-302)
};

// Ghost columns served straight from the NVLink mailbox of a row-partitioned level (comm.cu): the tile kernel waits --
// inside the kernel, right before its first tile that references ghost columns -- for the neighbours' flags of the
// current exchange and then gathers ghost entries from the mailbox itself.  No wait kernel, no copy into the ghost
// tail, no second launch for the boundary rows: one launch per operator application, like on one GPU.
struct GhostLink
{
  int rank;
  long long recv_off, recv_cnt; // position of this neighbour's entries in the ghost tail
  long long send_off, send_cnt, send_first; // entries this rank sends to it (index list offset / contiguous start)
};
struct GhostArgs
{
  int enabled = 0;
  int n_links = 0, nranks = 1;
  long long n_owned = 0;
  long long blo = 0, bhi = 0;          // rows [blo, bhi) reference owned columns only
  const unsigned char *local = nullptr; // this rank's peer window
  size_t box_off = 0, flag_off = 0;    // mailboxes [2][nranks][box_cap] doubles, flags [2][nranks] uint64
  long long box_cap = 0;
  const GhostLink *links = nullptr;    // device
  unsigned long long *seq = nullptr;   // device [2]: [1] = exchanges consumed so far
  unsigned int *done = nullptr;        // device [2]: [1] = consumer-warp completion counter
  unsigned long long timeout_ns = 0;
  int *err = nullptr;
  // push half of the exchange, done by the first n_push_ctas CTAs of the SAME kernel before they start on their tiles:
  // this rank's boundary entries of x -> the neighbours' mailboxes (remote stores), then the neighbours' flags
  int n_push_ctas = 0;
  int push_penalty = 0; // interior tiles a pushing CTA is spared (its stores + system fence cost about that much)
  int rank = 0;
  unsigned char *const *base = nullptr; // device: mapped windows of all ranks
  const int32_t *send_idx = nullptr;    // device: concatenated send lists (NULL: contiguous ranges)
};

struct EpiArgs
{
  double *y = nullptr;
  const double *b = nullptr;
  const double *dinv = nullptr;
  const double *xin = nullptr;
  double omega = 1.;
};

// one launch: y = epilogue(A x) on rows [row_begin, row_end) (row_end < 0: all rows)
int csr_apply(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args,
              int64_t row_begin = 0, int64_t row_end = -1);
// the direct-load (vector-CSR) kernels, whatever csr_uses_tile_kernel says
int csr_apply_vec(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args, int64_t row_begin,
                  int64_t row_end);
int choose_lanes(int64_t n_rows, int64_t nnz);
// tile-streamed variant (csr_tile.cu): same contract as csr_apply; requires A->tile_ok
int csr_apply_tile(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args, int64_t row_begin,
                   int64_t row_end, int64_t row_begin2 = 0, int64_t row_end2 = 0, const GhostArgs *ghost = nullptr);
// all rows of a partitioned level in ONE launch, ghost columns from the mailbox; false when the tile kernel cannot
// serve every row of A (then the caller takes the interior / wait / boundary path)
bool csr_can_fuse_ghost(const mfmgb_csr *A);
// two disjoint row ranges (the boundary blocks of a partitioned level): one launch with the tile kernel
int csr_apply2(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args, int64_t r0, int64_t r1,
               int64_t q0, int64_t q1);
int csr_measure_tiles(mfmgb_ctx *ctx, mfmgb_csr *A); // fills tile_cap (setup time, synchronises)
void csr_plan_tile(mfmgb_csr *A);                    // sets tile_ok / tile_stages / tile_ctas for the current lanes
int tile_cap_slot(int lanes);
bool csr_uses_tile_kernel(const mfmgb_csr *A);
} // namespace mfmgb

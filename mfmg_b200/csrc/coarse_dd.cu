// coarse_dd.cu -- domain-decomposed dense direct solve of the coarsest level of a row-partitioned hierarchy.
//
// The reference solves the coarsest level with one dense LU (source/cuda/cuda_solver.cu:496-515).  Replicating (or
// row-splitting) the dense inverse across the ranks reads 8 n_c^2 / N bytes per GPU and cycle, and in weak scaling
// n_c grows with N: the coarse solve ends up dominating the cycle.  The coarse operator of a z-slab partition is block
// tridiagonal in the agglomerate layers, so the SAME direct solve can be organised by one level of nested dissection:
//
//   coarse rows of rank r = [ interior I_r | separator S_r ]   (S_r = its last agglomerate layer; none on the last rank)
//   interiors of different ranks are coupled only through the separators, hence with S = union of the S_r
//
//     y_r   = A_II,r^-1 b_I,r                                     local dense GEMV      (8 n_I^2 bytes, constant in N)
//     t     = b_S - sum_r A_SI,r y_r                              ONE all-reduce of n_S doubles (the only communication)
//     x_S   = (A_SS - sum_r A_SI,r A_II,r^-1 A_IS,r)^-1 t         replicated dense GEMV (Schur complement, n_S = (N-1) L)
//     x_I,r = y_r - (A_II,r^-1 A_IS,r) x_S|adjacent               rectangular dense GEMV (n_I x 2L)
//
// This is block Gaussian elimination, i.e. the exact solve (rounding differs from the monolithic LU at the 1e-15 cond
// level).  Each rank ends up with x_c on its own rows and on ALL separators, which is what its prolongation rows read
// (own agglomerates + the layer below, a separator), so neither the right-hand side nor the solution is all-gathered.
// Setup (once): A_II^-1 by the dense factorisation of dense.cu, E_r = A_II^-1 A_IS and A_SI E_r by GEMM, the Schur
// matrix summed over the ranks by an all-reduce and inverted redundantly.
//
// Overlapped form (round 2, default over peer memory): steps 1-3 as written are a chain -- the all-reduce (latency +
// the skew between the ranks) and the Schur GEMV sit behind the interior GEMV.  With W_r = A_SI,r A_II,r^-1 formed
// once at setup (n_adj x n_I, dense) the separator right-hand side t = b_S - sum_r W_r b_I,r depends on the restricted
// residual only, so  [W b_I -> all-reduce -> Schur GEMV]  runs on a second stream NEXT TO  y = A_II^-1 b_I  and the
// two meet in the interior update.  Same block elimination, one more rounding-level reassociation ((A_SI A_II^-1) b
// instead of A_SI (A_II^-1 b)).
#include <algorithm>

#include "comm.cuh"
#include "dense.cuh"
#include "peer.cuh"

using namespace mfmgb;

struct mfmgb_coarse_dd
{
  int64_t n_c = 0;       // size of the coarse level
  int64_t own_begin = 0; // first coarse row of this rank; the interior is [own_begin, own_begin + n_I)
  int64_t n_I = 0, n_S = 0;
  int64_t adj_begin = 0, n_adj = 0;         // separators adjacent to this rank's interior, in S numbering
  int64_t own_sep_begin = 0, own_sep_n = 0; // this rank's own separator in S numbering (n = 0 on the last rank)
  mfmgb_dense *D_II = nullptr, *D_S = nullptr;
  const mfmgb_csr *A_SI = nullptr; // n_adj x n_I (borrowed)
  double *E = nullptr;             // n_I x ldE: A_II^-1 A_IS
  int64_t ldE = 0;
  int32_t *sep_index = nullptr; // [n_S] global coarse index of every separator row
  double *y = nullptr, *t = nullptr, *xs = nullptr, *g = nullptr;
  // overlapped form
  bool overlap = false;
  double *W = nullptr; // n_adj x ldW: A_SI A_II^-1
  int64_t ldW = 0;
  double *w = nullptr, *bI = nullptr; // W b_I; aligned copy of b_I when the caller's is not
  unsigned int *done = nullptr;       // CTA completion counter of dd_w_rhs_allreduce_kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace
{
struct DdRhsArgs
{
  int64_t n_S, adj_begin, n_adj, own_begin, own_n, n_below;
  const void *si_rowptr; // A_SI (n_adj x n_I), CSR
  const int *si_col;
  const double *si_val;
  const double *y;       // A_II^-1 b_I
  const int32_t *sep_index;
  const double *b_c, *g_below;
  double *t;
};

// this rank's contribution to the separator right-hand side, one WARP per entry (the row of A_SI is read by the 32
// lanes at once: a thread-per-row loop exposed one L2 round trip per non-zero, ~25 us for 27-entry rows):
// t[k] = (k adjacent ? -(A_SI y)[k - adj_begin] : 0) + (k in own separator ? b_c[sep_index[k]] : 0)
//        + (k in the separator below ? g_below[k - adj_begin] : 0)   [this rank's share of R r on those rows]
template <typename OffT>
__device__ __forceinline__ double dd_rhs_value_warp(const DdRhsArgs &a, int64_t k, int lane)
{
  double s = 0.;
  if (k >= a.adj_begin && k < a.adj_begin + a.n_adj)
  {
    const OffT *rp = static_cast<const OffT *>(a.si_rowptr);
    const int64_t r = k - a.adj_begin;
    const OffT je = rp[r + 1];
    for (OffT j = rp[r] + lane; j < je; j += 32)
      s = fma(a.si_val[j], a.y[a.si_col[j]], s);
  }
  s = subwarp_sum<32>(s); // fixed tree
  double v = -s;
  if (k >= a.own_begin && k < a.own_begin + a.own_n)
    v += a.b_c[a.sep_index[k]];
  if (a.g_below && k >= a.adj_begin && k < a.adj_begin + a.n_below)
    v += a.g_below[k - a.adj_begin];
  return v;
}

// ONE CTA: the local contributions, then the sum over the ranks through NVLink peer memory (peer.cuh) -- the SpMV with
// A_SI, the right-hand-side assembly and the all-reduce of the reference solve's Schur step in a single launch.
// A single CTA has little memory-level parallelism, so the SpMV is done in two flat passes instead of row by row (a
// warp-per-row loop chained ~3 L2 round trips per row, 8 rows per warp: ~20 us): every thread forms products
// val[j] * y[col[j]] for a strided share of ALL non-zeros into shared memory (two round trips in total), then one
// thread per row adds its segment in ascending order.
template <typename OffT>
__global__ void __launch_bounds__(1024) dd_rhs_allreduce_kernel(const DdRhsArgs a, const PeerAllreduceArgs pa, int prod_cap)
{
  extern __shared__ double dd_prod[];
  const OffT *rp = static_cast<const OffT *>(a.si_rowptr);
  const int64_t nnz = a.n_adj > 0 ? (int64_t)rp[a.n_adj] : 0;
  if (nnz <= prod_cap)
  {
    for (int64_t j = threadIdx.x; j < nnz; j += blockDim.x)
      dd_prod[j] = a.si_val[j] * a.y[a.si_col[j]];
    __syncthreads();
    for (int64_t k = threadIdx.x; k < a.n_S; k += blockDim.x)
    {
      double v = 0.;
      if (k >= a.adj_begin && k < a.adj_begin + a.n_adj)
      {
        const int64_t r = k - a.adj_begin;
        double s = 0.;
        for (OffT j = rp[r]; j < rp[r + 1]; ++j)
          s += dd_prod[j];
        v = -s;
      }
      if (k >= a.own_begin && k < a.own_begin + a.own_n)
        v += a.b_c[a.sep_index[k]];
      if (a.g_below && k >= a.adj_begin && k < a.adj_begin + a.n_below)
        v += a.g_below[k - a.adj_begin];
      a.t[k] = v;
    }
  }
  else
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    for (int64_t k = warp; k < a.n_S; k += n_warps)
    {
      const double v = dd_rhs_value_warp<OffT>(a, k, lane);
      if (lane == 0)
        a.t[k] = v;
    }
  }
  __syncthreads();
  peer_allreduce_cta(a.t, (int)a.n_S, pa);
}

// Overlapped form: w = W b_I with TPR threads per row (the loop of dense.cu's gemv_kernel), and the LAST CTA to finish
// assembles t = (own part of b_S) - w [+ share of R r below] and sums it over the ranks through peer memory -- the
// dense product, the right-hand-side assembly and the all-reduce in one launch.
template <int TPR>
__global__ void __launch_bounds__(256)
    dd_w_rhs_allreduce_kernel(int64_t n_I, const double *__restrict__ W, int64_t ldW, const double *__restrict__ bI,
                              double *__restrict__ w, const DdRhsArgs a, const PeerAllreduceArgs pa,
                              unsigned int *__restrict__ done)
{
  __shared__ double sm[8];
  __shared__ int is_last;
  constexpr int ROWS = 256 / TPR;
  const int t = threadIdx.x % TPR;
  const int64_t r = (int64_t)blockIdx.x * ROWS + threadIdx.x / TPR;
  double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
  if (r < a.n_adj)
  {
    const double2 *row2 = reinterpret_cast<const double2 *>(W + r * ldW);
    const double2 *v2 = reinterpret_cast<const double2 *>(bI);
    const int64_t npairs = n_I >> 1;
    int64_t q = t;
    for (; q + 3 * TPR < npairs; q += 4 * TPR)
    {
      const double2 a0 = ld_stream_f64x2(reinterpret_cast<const double *>(row2 + q));
      const double2 a1 = ld_stream_f64x2(reinterpret_cast<const double *>(row2 + q + TPR));
      const double2 a2 = ld_stream_f64x2(reinterpret_cast<const double *>(row2 + q + 2 * TPR));
      const double2 a3 = ld_stream_f64x2(reinterpret_cast<const double *>(row2 + q + 3 * TPR));
      const double2 b0 = v2[q], b1 = v2[q + TPR], b2 = v2[q + 2 * TPR], b3 = v2[q + 3 * TPR];
      s0 = fma(a0.y, b0.y, fma(a0.x, b0.x, s0));
      s1 = fma(a1.y, b1.y, fma(a1.x, b1.x, s1));
      s2 = fma(a2.y, b2.y, fma(a2.x, b2.x, s2));
      s3 = fma(a3.y, b3.y, fma(a3.x, b3.x, s3));
    }
    for (; q < npairs; q += TPR)
    {
      const double2 a0 = ld_stream_f64x2(reinterpret_cast<const double *>(row2 + q));
      const double2 b0 = v2[q];
      s0 = fma(a0.y, b0.y, fma(a0.x, b0.x, s0));
    }
    if ((n_I & 1) && t == 0)
      s1 = fma(W[r * ldW + n_I - 1], bI[n_I - 1], s1);
  }
  double s = (s0 + s1) + (s2 + s3);
  if (TPR == 256)
  {
    s = block_sum<256>(s, sm);
    if (threadIdx.x == 0 && r < a.n_adj)
      w[r] = s;
  }
  else
  {
    s = subwarp_sum<32>(s);
    if (t == 0 && r < a.n_adj)
      w[r] = s;
  }
  // completion count: the writers fence, the barrier orders them before thread 0's ticket
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    is_last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!is_last)
    return;
  if (threadIdx.x == 0)
    *done = 0; // for the next launch (graph replay)
  __threadfence();
  for (int64_t k = threadIdx.x; k < a.n_S; k += blockDim.x)
  {
    double v = 0.;
    if (k >= a.adj_begin && k < a.adj_begin + a.n_adj)
      v = -__ldcg(w + (k - a.adj_begin));
    if (k >= a.own_begin && k < a.own_begin + a.own_n)
      v += a.b_c[a.sep_index[k]];
    if (a.g_below && k >= a.adj_begin && k < a.adj_begin + a.n_below)
      v += a.g_below[k - a.adj_begin];
    a.t[k] = v;
  }
  __syncthreads();
  peer_allreduce_cta(a.t, (int)a.n_S, pa);
}

// the same without the reduction (NCCL transport: ncclAllReduce follows)
template <typename OffT>
__global__ void __launch_bounds__(256) dd_rhs_kernel(const DdRhsArgs a)
{
  const int64_t k = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k < a.n_S)
  {
    const double v = dd_rhs_value_warp<OffT>(a, k, lane);
    if (lane == 0)
      a.t[k] = v;
  }
}

// blocks [0, interior_blocks): x_I[i] = y[i] - sum_j E[i][j] xs[j]   (one warp per row, fixed shuffle tree)
// the remaining blocks:        x_c[sep_index[k]] = xs_all[k]         (the separator solution, replicated)
__global__ void __launch_bounds__(256)
    dd_interior_scatter_kernel(int64_t n_I, int64_t n_adj, const double *__restrict__ E, int64_t ldE,
                               const double *__restrict__ xs, const double *__restrict__ y, double *__restrict__ x_I,
                               int interior_blocks, int64_t n_S, const int32_t *__restrict__ sep_index,
                               const double *__restrict__ xs_all, double *__restrict__ x_c)
{
  if ((int)blockIdx.x >= interior_blocks)
  {
    const int64_t k = (int64_t)(blockIdx.x - interior_blocks) * 256 + threadIdx.x;
    if (k < n_S)
      x_c[sep_index[k]] = xs_all[k];
    return;
  }
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double s0 = 0., s1 = 0.;
  if (i < n_I)
  {
    const double *row = E + i * ldE;
    int64_t j = lane;
    for (; j + 32 < n_adj; j += 64)
    {
      s0 = fma(row[j], xs[j], s0);
      s1 = fma(row[j + 32], xs[j + 32], s1);
    }
    if (j < n_adj)
      s0 = fma(row[j], xs[j], s0);
  }
  const double s = subwarp_sum<32>(s0 + s1);
  if (lane == 0 && i < n_I)
    x_I[i] = y[i] - s;
}

// S[(adj_begin + i) * ldS + adj_begin + j] -= C[i * ldC + j]
__global__ void __launch_bounds__(256) dd_sub_block_kernel(int64_t n_adj, const double *__restrict__ C, int64_t ldC,
                                                           double *__restrict__ S, int64_t ldS, int64_t adj_begin)
{
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j < n_adj)
    S[(adj_begin + i) * ldS + adj_begin + j] -= C[i * ldC + j];
}
} // namespace

namespace mfmgb
{
int64_t coarse_dd_n_sep_below(const mfmgb_coarse_dd *d) { return d->n_adj - d->own_sep_n; }

int coarse_dd_solve_async(mfmgb_ctx *ctx, const mfmgb_coarse_dd *d, const double *b_c, double *x_c,
                          const double *g_below)
{
  mfmgb_comm *c = ctx_comm(ctx);
  cudaStream_t st = ctx->stream;
  if (d->overlap && d->n_S > 0 && c->peer.enabled && d->n_S <= c->peer.ar_cap)
  {
    // [W b_I -> all-reduce -> Schur GEMV] on the communication stream next to y = A_II^-1 b_I on the compute stream
    const double *bI = b_c + d->own_begin;
    if (d->n_I > 0 && (reinterpret_cast<uintptr_t>(bI) & 15))
    {
      MFMGB_CUDA(ctx, cudaMemcpyAsync(d->bI, bI, sizeof(double) * (size_t)d->n_I, cudaMemcpyDeviceToDevice, st));
      bI = d->bI;
    }
    DdRhsArgs a;
    a.n_S = d->n_S;
    a.adj_begin = d->adj_begin;
    a.n_adj = d->n_I > 0 ? d->n_adj : 0;
    a.own_begin = d->own_sep_begin;
    a.own_n = d->own_sep_n;
    a.n_below = d->n_adj - d->own_sep_n;
    a.si_rowptr = nullptr;
    a.si_col = nullptr;
    a.si_val = nullptr;
    a.y = nullptr;
    a.sep_index = d->sep_index;
    a.b_c = b_c;
    a.g_below = g_below;
    a.t = d->t;
    MFMGB_CUDA(ctx, cudaEventRecord(d->ev_fork, st));
    MFMGB_CUDA(ctx, cudaStreamWaitEvent(c->stream, d->ev_fork, 0));
    if (d->n_I >= 1024)
      dd_w_rhs_allreduce_kernel<256><<<(unsigned)std::max<int64_t>(a.n_adj, 1), 256, 0, c->stream>>>(
          d->n_I, d->W, d->ldW, bI, d->w, a, peer_allreduce_args(c), d->done);
    else
      dd_w_rhs_allreduce_kernel<32><<<(unsigned)std::max<int64_t>(ceil_div(a.n_adj, 8), 1), 256, 0, c->stream>>>(
          d->n_I, d->W, d->ldW, bI, d->w, a, peer_allreduce_args(c), d->done);
    MFMGB_LAUNCHED(ctx);
    ctx->stream = c->stream; // the Schur solve follows on the same stream
    const int rc = dense_solve_async(ctx, d->D_S, d->t, d->xs);
    ctx->stream = st;
    MFMGB_CHECK(rc);
    MFMGB_CUDA(ctx, cudaEventRecord(d->ev_join, c->stream));
    if (d->n_I > 0)
      MFMGB_CHECK(dense_solve_async(ctx, d->D_II, bI, d->y));
    MFMGB_CUDA(ctx, cudaStreamWaitEvent(st, d->ev_join, 0));
    prof_mark(ctx, "coarse: y = A_II^-1 b_I  ||  W b_I + all-reduce + Schur GEMV");
    const int interior_blocks = (int)ceil_div(d->n_I, 8);
    const int scatter_blocks = (int)ceil_div(d->n_S, 256);
    dd_interior_scatter_kernel<<<(unsigned)(interior_blocks + scatter_blocks), 256, 0, st>>>(
        d->n_I, d->n_adj, d->E, d->ldE, d->xs + d->adj_begin, d->y, x_c + d->own_begin, interior_blocks, d->n_S,
        d->sep_index, d->xs, x_c);
    MFMGB_LAUNCHED(ctx);
    prof_mark(ctx, "coarse: interior update + scatter");
    return MFMGB_OK;
  }
  // y = A_II^-1 b_I
  if (d->n_I > 0)
    MFMGB_CHECK(dense_solve_async(ctx, d->D_II, b_c + d->own_begin, d->y));
  prof_mark(ctx, "coarse: y = A_II^-1 b_I (GEMV)");
  if (d->n_S > 0)
  {
    // t = (own part of b_S) - A_SI y [+ this rank's share of R r below], summed over the ranks
    DdRhsArgs a;
    a.n_S = d->n_S;
    a.adj_begin = d->adj_begin;
    a.n_adj = d->n_I > 0 ? d->n_adj : 0;
    a.own_begin = d->own_sep_begin;
    a.own_n = d->own_sep_n;
    a.n_below = d->n_adj - d->own_sep_n;
    a.si_rowptr = d->A_SI->rowptr;
    a.si_col = d->A_SI->col;
    a.si_val = d->A_SI->val;
    a.y = d->y;
    a.sep_index = d->sep_index;
    a.b_c = b_c;
    a.g_below = g_below;
    a.t = d->t;
    if (c->peer.enabled && d->n_S <= c->peer.ar_cap)
    {
      // one launch: contributions + sum over the ranks through peer memory
      // products of all non-zeros of A_SI in shared memory when they fit (they do: a few thousand entries)
      const int prod_cap = (int)std::min<int64_t>(std::max<int64_t>(d->A_SI->nnz, 1), 20000);
      const size_t smem = sizeof(double) * (size_t)prod_cap;
      static unsigned long long configured = 0; // one bit per device
      if (!((configured >> (ctx->device & 63)) & 1ull))
      {
        MFMGB_CUDA(ctx, cudaFuncSetAttribute(dd_rhs_allreduce_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(160 * 1024)));
        MFMGB_CUDA(ctx, cudaFuncSetAttribute(dd_rhs_allreduce_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(160 * 1024)));
        configured |= 1ull << (ctx->device & 63);
      }
      if (d->A_SI->off64)
        dd_rhs_allreduce_kernel<int64_t><<<1, 1024, smem, st>>>(a, peer_allreduce_args(c), prod_cap);
      else
        dd_rhs_allreduce_kernel<int32_t><<<1, 1024, smem, st>>>(a, peer_allreduce_args(c), prod_cap);
      MFMGB_LAUNCHED(ctx);
    }
    else
    {
      if (d->A_SI->off64)
        dd_rhs_kernel<int64_t><<<(unsigned)ceil_div(d->n_S, 8), 256, 0, st>>>(a);
      else
        dd_rhs_kernel<int32_t><<<(unsigned)ceil_div(d->n_S, 8), 256, 0, st>>>(a);
      MFMGB_LAUNCHED(ctx);
      MFMGB_CHECK(allreduce_sum(ctx, d->t, (int)d->n_S));
    }
    prof_mark(ctx, "coarse: separator rhs + all-reduce");
    // x_S = Schur^-1 t (replicated)
    MFMGB_CHECK(dense_solve_async(ctx, d->D_S, d->t, d->xs));
    prof_mark(ctx, "coarse: x_S = Schur^-1 t (GEMV)");
  }
  // x_I = y - E x_S|adjacent, and the separator solution into x_c: one launch
  const int interior_blocks = (int)ceil_div(d->n_I, 8);
  const int scatter_blocks = (int)ceil_div(d->n_S, 256);
  if (interior_blocks + scatter_blocks > 0)
  {
    dd_interior_scatter_kernel<<<(unsigned)(interior_blocks + scatter_blocks), 256, 0, st>>>(
        d->n_I, d->n_adj, d->E, d->ldE, d->xs + d->adj_begin, d->y, x_c + d->own_begin, interior_blocks, d->n_S,
        d->sep_index, d->xs, x_c);
    MFMGB_LAUNCHED(ctx);
  }
  prof_mark(ctx, "coarse: interior update + scatter");
  return MFMGB_OK;
}
} // namespace mfmgb

extern "C"
{
  MFMGB_API int mfmgb_coarse_dd_create(mfmgb_ctx *ctx, int64_t n_c, int64_t own_begin, int64_t n_S, int64_t adj_begin,
                                       int64_t own_sep_begin, int64_t own_sep_n, const mfmgb_csr *A_II,
                                       const mfmgb_csr *A_IS, const mfmgb_csr *A_SI, const mfmgb_csr *A_SS,
                                       const int32_t *sep_index, mfmgb_coarse_dd **out)
  {
    NvtxRange nvtx_range("mfmgb: coarse solver setup (domain-decomposed)");
    MFMGB_REQUIRE(ctx, ctx && out && A_II && A_IS && A_SI && A_SS && (n_S == 0 || sep_index),
                  "mfmgb_coarse_dd_create: bad arguments");
    mfmgb_comm *c = ctx_comm(ctx);
    MFMGB_REQUIRE(ctx, c != nullptr, "mfmgb_coarse_dd_create: needs an initialised communicator (mfmgb_comm_init)");
    const int64_t n_I = A_II->n_rows, n_adj = A_IS->n_cols;
    MFMGB_REQUIRE(ctx, A_II->n_cols == n_I && A_IS->n_rows == n_I && A_SI->n_rows == n_adj && A_SI->n_cols == n_I &&
                           A_SS->n_rows == n_S && A_SS->n_cols == n_S,
                  "mfmgb_coarse_dd_create: block shapes are inconsistent");
    MFMGB_REQUIRE(ctx, own_begin >= 0 && own_begin + n_I <= n_c && adj_begin >= 0 && adj_begin + n_adj <= n_S &&
                           own_sep_begin >= 0 && own_sep_begin + own_sep_n <= n_S,
                  "mfmgb_coarse_dd_create: index ranges out of bounds");
    *out = nullptr;
    mfmgb_coarse_dd *d = new mfmgb_coarse_dd();
    d->n_c = n_c;
    d->own_begin = own_begin;
    d->n_I = n_I;
    d->n_S = n_S;
    d->adj_begin = adj_begin;
    d->n_adj = n_adj;
    d->own_sep_begin = own_sep_begin;
    d->own_sep_n = own_sep_n;
    d->A_SI = A_SI;
    cudaStream_t st = ctx->stream;
    const int64_t ldE = (std::max<int64_t>(n_adj, 1) + 3) & ~(int64_t)3;
    d->ldE = ldE;
    MFMGB_CUDA(ctx, cudaMalloc(&d->y, sizeof(double) * (size_t)(n_I + 2)));
    MFMGB_CUDA(ctx, cudaMalloc(&d->t, sizeof(double) * (size_t)(n_S + 2)));
    MFMGB_CUDA(ctx, cudaMalloc(&d->xs, sizeof(double) * (size_t)(n_S + 2)));
    MFMGB_CUDA(ctx, cudaMalloc(&d->g, sizeof(double) * (size_t)(n_adj + 2)));
    MFMGB_CUDA(ctx, cudaMemsetAsync(d->g, 0, sizeof(double) * (size_t)(n_adj + 2), st));
    MFMGB_CUDA(ctx, cudaMemsetAsync(d->xs, 0, sizeof(double) * (size_t)(n_S + 2), st));
    if (n_S > 0)
    {
      MFMGB_CUDA(ctx, cudaMalloc(&d->sep_index, sizeof(int32_t) * (size_t)n_S));
      MFMGB_CUDA(ctx, cudaMemcpy(d->sep_index, sep_index, sizeof(int32_t) * (size_t)n_S, cudaMemcpyHostToDevice));
    }
    // interior block: A_II^-1
    if (n_I > 0)
      MFMGB_CHECK(mfmgb_dense_factor(ctx, A_II, &d->D_II));
    // overlapped form: over peer memory, unless MFMGB_COARSE_OVERLAP=0 (measurement aid)
    const char *ov = getenv("MFMGB_COARSE_OVERLAP");
    const bool want_overlap = c->peer.enabled && n_S > 0 && n_S <= c->peer.ar_cap && !(ov && ov[0] == '0');
    if (n_S > 0)
    {
      const int64_t ldS = (n_S + 3) & ~(int64_t)3;
      double *S = nullptr;
      // rank 0 contributes A_SS, every rank subtracts its A_SI A_II^-1 A_IS block; the sum is the Schur complement
      if (c->rank == 0)
        MFMGB_CHECK(csr_to_dense_device(ctx, A_SS, ldS, &S));
      else
      {
        MFMGB_CUDA(ctx, cudaMalloc(&S, sizeof(double) * (size_t)(n_S * ldS)));
        MFMGB_CUDA(ctx, cudaMemsetAsync(S, 0, sizeof(double) * (size_t)(n_S * ldS), st));
      }
      if (n_I > 0 && n_adj > 0)
      {
        double *Ais = nullptr, *Asi = nullptr, *C = nullptr;
        const int64_t ldI = (n_I + 3) & ~(int64_t)3;
        MFMGB_CHECK(csr_to_dense_device(ctx, A_IS, ldE, &Ais)); // n_I x ldE
        MFMGB_CHECK(csr_to_dense_device(ctx, A_SI, ldI, &Asi)); // n_adj x ldI
        MFMGB_CUDA(ctx, cudaMalloc(&d->E, sizeof(double) * (size_t)(n_I * ldE)));
        MFMGB_CUDA(ctx, cudaMemsetAsync(d->E, 0, sizeof(double) * (size_t)(n_I * ldE), st));
        MFMGB_CUDA(ctx, cudaMalloc(&C, sizeof(double) * (size_t)(n_adj * ldE)));
        // E = A_II^-1 A_IS ; C = A_SI E
        MFMGB_CHECK(dense_gemm(ctx, n_I, n_adj, n_I, d->D_II->inv, d->D_II->lda, Ais, ldE, d->E, ldE, 1., 0.));
        MFMGB_CHECK(dense_gemm(ctx, n_adj, n_adj, n_I, Asi, ldI, d->E, ldE, C, ldE, 1., 0.));
        dim3 grid((unsigned)ceil_div(n_adj, 256), (unsigned)n_adj);
        dd_sub_block_kernel<<<grid, 256, 0, st>>>(n_adj, C, ldE, S, ldS, adj_begin);
        MFMGB_LAUNCHED(ctx);
        if (want_overlap && !d->D_II->substitution)
        {
          // W = A_SI A_II^-1
          d->ldW = ldI;
          MFMGB_CUDA(ctx, cudaMalloc(&d->W, sizeof(double) * (size_t)(n_adj * ldI)));
          MFMGB_CUDA(ctx, cudaMemsetAsync(d->W, 0, sizeof(double) * (size_t)(n_adj * ldI), st));
          MFMGB_CHECK(dense_gemm(ctx, n_adj, n_I, n_I, Asi, ldI, d->D_II->inv, d->D_II->lda, d->W, ldI, 1., 0.));
        }
        MFMGB_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(Ais);
        cudaFree(Asi);
        cudaFree(C);
      }
      MFMGB_NCCL(ctx, ncclAllReduce(S, S, (size_t)(n_S * ldS), ncclDouble, ncclSum, c->nccl, st));
      MFMGB_CHECK(dense_factor_device(ctx, S, n_S, &d->D_S)); // takes ownership of S
      // every rank must take the same branch (each form issues exactly one peer all-reduce, so mixing them would
      // still pair up, but keep the cycle symmetric): overlap unless some rank had to keep its interior factors
      long long veto = (want_overlap && (n_I == 0 || n_adj == 0 || d->W != nullptr)) ? 0 : 1;
      MFMGB_CHECK(comm_agree_max(ctx, &veto));
      d->overlap = veto == 0;
      if (d->overlap)
      {
        d->D_S->direct_gemv = true; // runs next to the streamed interior GEMV: no shared-memory footprint
        MFMGB_CUDA(ctx, cudaMalloc(&d->w, sizeof(double) * (size_t)(n_adj + 2)));
        MFMGB_CUDA(ctx, cudaMemsetAsync(d->w, 0, sizeof(double) * (size_t)(n_adj + 2), st));
        MFMGB_CUDA(ctx, cudaMalloc(&d->bI, sizeof(double) * (size_t)(n_I + 2)));
        MFMGB_CUDA(ctx, cudaMalloc(&d->done, sizeof(unsigned int)));
        MFMGB_CUDA(ctx, cudaMemsetAsync(d->done, 0, sizeof(unsigned int), st));
        MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
        MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
      }
    }
    MFMGB_CUDA(ctx, cudaStreamSynchronize(st));
    *out = d;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_coarse_dd_destroy(mfmgb_ctx *ctx, mfmgb_coarse_dd *d)
  {
    if (!d)
      return MFMGB_OK;
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    mfmgb_dense_destroy(ctx, d->D_II);
    mfmgb_dense_destroy(ctx, d->D_S);
    cudaFree(d->E);
    cudaFree(d->W);
    cudaFree(d->w);
    cudaFree(d->bI);
    cudaFree(d->done);
    if (d->ev_fork)
      cudaEventDestroy(d->ev_fork);
    if (d->ev_join)
      cudaEventDestroy(d->ev_join);
    cudaFree(d->sep_index);
    cudaFree(d->y);
    cudaFree(d->t);
    cudaFree(d->xs);
    cudaFree(d->g);
    delete d;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_coarse_dd_solve(mfmgb_ctx *ctx, const mfmgb_coarse_dd *d, const double *b_c, double *x_c)
  {
    MFMGB_REQUIRE(ctx, ctx && d && b_c && x_c && b_c != x_c, "mfmgb_coarse_dd_solve: bad arguments");
    return coarse_dd_solve_async(ctx, d, b_c, x_c, nullptr);
  }
}

// csr.cu -- CSR upload/transposition and the vector-CSR SpMV kernels with fused epilogues.
//
// Kernel: LPR lanes (a power of two, 2..32) cooperate on one row; LPR is chosen per matrix
// (i.e. per level) from the mean row length.  Column indices and values are streamed once,
// x is gathered through the read-only path and stays L2/L1 resident; the row sum is reduced
// with a fixed shuffle tree (no atomics => deterministic) and the epilogue (residual, Jacobi
// update, prolongation correction) is applied in the same kernel, so no intermediate vector
// ever goes to HBM.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "csr.cuh"

using namespace mfmgb;

namespace mfmgb
{
namespace
{
constexpr int kBlock = 256;
constexpr size_t kRowptrPad = 264; // >= rows per tile (256) + 4, see csr_tile.cu

// epilogue operands of one row; requested before the dot-product loop so that their latency overlaps it
struct EpiRegs
{
  double b = 0., dinv = 0., x = 0.;
};

template <int EPI>
__device__ __forceinline__ EpiRegs epilogue_load(const EpiArgs &e, int64_t row)
{
  EpiRegs r;
  if (EPI == (int)Epi::Resid || EPI == (int)Epi::Jacobi)
    r.b = e.b[row];
  if (EPI == (int)Epi::Jacobi)
  {
    r.dinv = e.dinv[row];
    r.x = e.xin[row];
  }
  if (EPI == (int)Epi::Sub)
    r.x = e.y[row];
  return r;
}

template <int EPI>
__device__ __forceinline__ void epilogue(const EpiArgs &e, const EpiRegs &g, int64_t row, double s)
{
  if (EPI == (int)Epi::Spmv)
    e.y[row] = s;
  else if (EPI == (int)Epi::Resid)
    e.y[row] = __dsub_rn(s, g.b);
  else if (EPI == (int)Epi::Jacobi)
  {
    const double r = __dsub_rn(s, g.b);
    double t = __dmul_rn(g.dinv, r);
    if (e.omega != 1.)
      t = __dmul_rn(e.omega, t);
    e.y[row] = __dsub_rn(g.x, t);
  }
  else
    e.y[row] = __dsub_rn(g.x, s);
}

template <int LPR, int EPI, typename OffT>
__global__ void __launch_bounds__(kBlock)
    csr_vec_kernel(int64_t row_begin, int64_t n_rows, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                   const double *__restrict__ val, const double *__restrict__ x, EpiArgs e)
{
  // rows [row_begin, n_rows) are processed by this launch
  const int64_t row = row_begin + ((int64_t)blockIdx.x * kBlock + threadIdx.x) / LPR;
  const int lane = threadIdx.x & (LPR - 1);
  double s0 = 0., s1 = 0.;
  EpiRegs g;
  if (row < n_rows)
  {
    OffT k = rowptr[row] + lane;
    const OffT k1 = rowptr[row + 1];
    if (lane == 0)
      g = epilogue_load<EPI>(e, row);
    // two independent (col, val, x) chains per lane; a 4-way unroll measured 13 % slower on B200
    // (profiles/r01_lanes_sweep.md)
    for (; k + LPR < k1; k += 2 * LPR)
    {
      const int c0 = col[k], c1 = col[k + LPR];
      const double v0 = val[k], v1 = val[k + LPR];
      s0 = fma(v0, __ldg(x + c0), s0);
      s1 = fma(v1, __ldg(x + c1), s1);
    }
    if (k < k1)
      s0 = fma(val[k], __ldg(x + col[k]), s0);
  }
  double s = subwarp_sum<LPR>(s0 + s1);
  if (lane == 0 && row < n_rows)
    epilogue<EPI>(e, g, row, s);
}

// one THREAD per row, RPT rows per thread in flight: operators with one or two entries per row (the prolongation P = R^T:
// 1.2 - 1.4 entries per row).  With one row per thread the kernel is a chain of three dependent loads (row offsets ->
// column / value -> x gather) and runs at occupancy x latency: 17 M rows in 0.133 ms = 4.4 TB/s at cfg4.  Here the loads
// of RPT rows are issued together, unconditionally (clamped indices; a predicated load is paired with its use by ptxas,
// DESIGN.md section 4), so RPT chains overlap.  Summation order of csr_vec_kernel<1> (even entries -> s0, odd -> s1):
// identical bits.
template <int EPI, typename OffT, int RPT>
__global__ void __launch_bounds__(kBlock)
    csr_short_kernel(int64_t row_begin, int64_t n_rows, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                     const double *__restrict__ val, const double *__restrict__ x, EpiArgs e)
{
  const int64_t base = row_begin + (int64_t)blockIdx.x * (kBlock * RPT) + threadIdx.x;
  OffT k0[RPT], k1[RPT];
  EpiRegs g[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j)
  {
    const int64_t row = base + (int64_t)j * kBlock;
    const int64_t rc = row < n_rows ? row : n_rows - 1;
    k0[j] = rowptr[rc];
    k1[j] = rowptr[rc + 1];
    g[j] = epilogue_load<EPI>(e, rc);
  }
  int c[RPT];
  double v[RPT], xv[RPT], s0[RPT], s1[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j)
  {
    const OffT k = k0[j] < k1[j] ? k0[j] : (OffT)0; // empty row: entry 0 of the matrix (read, not used)
    c[j] = col[k];
    v[j] = val[k];
  }
#pragma unroll
  for (int j = 0; j < RPT; ++j)
    xv[j] = __ldg(x + c[j]);
#pragma unroll
  for (int j = 0; j < RPT; ++j)
  {
    s0[j] = k0[j] < k1[j] ? fma(v[j], xv[j], 0.) : 0.;
    s1[j] = 0.;
  }
#pragma unroll
  for (int j = 0; j < RPT; ++j)
    for (OffT k = k0[j] + 1; k < k1[j]; ++k)
    {
      const double p = val[k], q = __ldg(x + col[k]);
      if ((k - k0[j]) & 1)
        s1[j] = fma(p, q, s1[j]);
      else
        s0[j] = fma(p, q, s0[j]);
    }
#pragma unroll
  for (int j = 0; j < RPT; ++j)
  {
    const int64_t row = base + (int64_t)j * kBlock;
    if (row < n_rows)
      epilogue<EPI>(e, g[j], row, s0[j] + s1[j]);
  }
}

// one CTA per row: rows of thousands of entries (restrictors of large agglomerates: 17^3 = 4913, 21^3 = 9261) in
// matrices with too few rows to fill the GPU with one warp per row.  4 gathers in flight per thread, fixed block tree.
template <int EPI, typename OffT>
__global__ void __launch_bounds__(kBlock, 4)
    csr_rowblock_kernel(int64_t row_begin, int64_t n_rows, const OffT *__restrict__ rowptr, const int *__restrict__ col,
                        const double *__restrict__ val, const double *__restrict__ x, EpiArgs e)
{
  __shared__ double sm[kBlock / 32];
  const int64_t row = row_begin + blockIdx.x;
  if (row >= n_rows)
    return;
  EpiRegs g;
  if (threadIdx.x == 0)
    g = epilogue_load<EPI>(e, row);
  const OffT k1 = rowptr[row + 1];
  OffT k = rowptr[row] + threadIdx.x;
  double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
  for (; k + 3 * kBlock < k1; k += 4 * kBlock)
  {
    const int c0 = col[k], c1 = col[k + kBlock], c2 = col[k + 2 * kBlock], c3 = col[k + 3 * kBlock];
    const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
    s0 = fma(val[k], x0, s0);
    s1 = fma(val[k + kBlock], x1, s1);
    s2 = fma(val[k + 2 * kBlock], x2, s2);
    s3 = fma(val[k + 3 * kBlock], x3, s3);
  }
  for (; k < k1; k += kBlock)
    s0 = fma(val[k], __ldg(x + col[k]), s0);
  const double s = block_sum<kBlock>((s0 + s1) + (s2 + s3), sm);
  if (threadIdx.x == 0)
    epilogue<EPI>(e, g, row, s);
}

template <int EPI, typename OffT>
int launch_rowblock(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const EpiArgs &e, int64_t r0, int64_t r1)
{
  if (r1 <= r0)
    return MFMGB_OK;
  csr_rowblock_kernel<EPI, OffT><<<(unsigned)(r1 - r0), kBlock, 0, ctx->stream>>>(r0, r1, (const OffT *)A->rowptr, A->col,
                                                                                 A->val, x, e);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

template <int LPR, int EPI, typename OffT>
int launch_vec(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const EpiArgs &e, int64_t r0, int64_t r1)
{
  const int64_t threads = (r1 - r0) * LPR;
  const int64_t nb = ceil_div(threads, kBlock);
  if (nb <= 0)
    return MFMGB_OK;
  csr_vec_kernel<LPR, EPI, OffT><<<(unsigned)nb, kBlock, 0, ctx->stream>>>(
      r0, r1, (const OffT *)A->rowptr, A->col, A->val, x, e);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

template <int EPI, typename OffT>
int launch_short(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const EpiArgs &e, int64_t r0, int64_t r1)
{
  constexpr int RPT = 4;
  const int64_t nb = ceil_div(r1 - r0, (int64_t)kBlock * RPT);
  if (nb <= 0)
    return MFMGB_OK;
  csr_short_kernel<EPI, OffT, RPT><<<(unsigned)nb, kBlock, 0, ctx->stream>>>(r0, r1, (const OffT *)A->rowptr, A->col,
                                                                             A->val, x, e);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

template <int EPI, typename OffT>
int dispatch_lanes(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const EpiArgs &e, int64_t r0, int64_t r1)
{
  // MFMGB_CSR_SHORT=0: one row per thread for lanes == 1 (measurement aid)
  static const bool short_rows = [] {
    const char *v = getenv("MFMGB_CSR_SHORT");
    return !(v && v[0] == '0');
  }();
  switch (A->lanes)
  {
  case 1:
    if (short_rows && A->nnz > 0 && r1 - r0 >= 4096)
      return launch_short<EPI, OffT>(ctx, A, x, e, r0, r1);
    return launch_vec<1, EPI, OffT>(ctx, A, x, e, r0, r1);
  case 2:
    return launch_vec<2, EPI, OffT>(ctx, A, x, e, r0, r1);
  case 4:
    return launch_vec<4, EPI, OffT>(ctx, A, x, e, r0, r1);
  case 8:
    return launch_vec<8, EPI, OffT>(ctx, A, x, e, r0, r1);
  case 16:
    return launch_vec<16, EPI, OffT>(ctx, A, x, e, r0, r1);
  case 256:
    return launch_rowblock<EPI, OffT>(ctx, A, x, e, r0, r1);
  default:
    return launch_vec<32, EPI, OffT>(ctx, A, x, e, r0, r1);
  }
}

template <typename OffT>
int dispatch_epi(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &e, int64_t r0,
                 int64_t r1)
{
  switch (epi)
  {
  case Epi::Spmv:
    return dispatch_lanes<(int)Epi::Spmv, OffT>(ctx, A, x, e, r0, r1);
  case Epi::Resid:
    return dispatch_lanes<(int)Epi::Resid, OffT>(ctx, A, x, e, r0, r1);
  case Epi::Jacobi:
    return dispatch_lanes<(int)Epi::Jacobi, OffT>(ctx, A, x, e, r0, r1);
  default:
    return dispatch_lanes<(int)Epi::Sub, OffT>(ctx, A, x, e, r0, r1);
  }
}
} // namespace

int choose_lanes(int64_t n_rows, int64_t nnz)
{
  if (n_rows <= 0)
    return 8;
  // Measured on B200 (profiles/r01_lanes_sweep.md): about 6-7 entries per lane is the sweet spot
  // (27-nnz rows: 4 lanes reach 88 % of the copy bandwidth, 8 lanes 65 %, 32 lanes 28 %).
  const double mean = (double)nnz / (double)n_rows;
  // thousands of entries per row and too few rows for one warp each to fill 148 SMs: one CTA per row
  if (mean >= 2048. && n_rows < 16384)
    return 256;
  int lanes = 1;
  while (lanes < 32 && lanes * 2 * 6 <= mean)
    lanes *= 2;
  return lanes;
}

bool csr_uses_tile_kernel(const mfmgb_csr *A)
{
  if (!A->tile_ok || A->kernel_override == 0)
    return false;
  if (A->kernel_override == 1)
    return true;
  static const int env_kernel = [] {
    const char *v = getenv("MFMGB_CSR_KERNEL");
    return !v ? -1 : (!strcmp(v, "vec") ? 0 : (!strcmp(v, "tile") ? 1 : -1));
  }();
  if (env_kernel >= 0)
    return env_kernel == 1;
  // automatic: the TMA ring pays off on long streams; tiny operators stay on the direct-load kernel
  // (MFMGB_TILE_MIN_ROWS / MFMGB_TILE_MIN_ROW_NNZ move the thresholds: measurement aid)
  static const int64_t min_rows = [] {
    const char *v = getenv("MFMGB_TILE_MIN_ROWS");
    return v && *v ? atoll(v) : (int64_t)32768;
  }();
  static const int64_t min_row_nnz = [] {
    const char *v = getenv("MFMGB_TILE_MIN_ROW_NNZ");
    return v && *v ? atoll(v) : (int64_t)8;
  }();
  return A->n_rows >= min_rows && A->nnz >= min_row_nnz * A->n_rows;
}

int csr_apply(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args, int64_t row_begin,
              int64_t row_end)
{
  if (A->n_rows >= ((int64_t)1 << 31) * 8)
    return fail(ctx, MFMGB_ERR_INVALID, "csr_apply: too many rows");
  if (row_end < 0)
    row_end = A->n_rows;
  if (row_begin < 0 || row_end > A->n_rows)
    return fail(ctx, MFMGB_ERR_INVALID, "csr_apply: row range out of bounds");
  if (csr_uses_tile_kernel(A))
    return csr_apply_tile(ctx, A, x, epi, args, row_begin, row_end);
  return csr_apply_vec(ctx, A, x, epi, args, row_begin, row_end);
}

int csr_apply_vec(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args, int64_t row_begin,
                  int64_t row_end)
{
  if (A->off64)
    return dispatch_epi<int64_t>(ctx, A, x, epi, args, row_begin, row_end);
  return dispatch_epi<int32_t>(ctx, A, x, epi, args, row_begin, row_end);
}

int csr_apply2(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args, int64_t r0, int64_t r1,
               int64_t q0, int64_t q1)
{
  if (r0 < 0 || r1 > A->n_rows || q0 < 0 || q1 > A->n_rows || (r1 > r0 && q1 > q0 && q0 < r1))
    return fail(ctx, MFMGB_ERR_INVALID, "csr_apply2: row ranges out of bounds or overlapping");
  if (csr_uses_tile_kernel(A))
    return csr_apply_tile(ctx, A, x, epi, args, r0, r1, q0, q1);
  if (r1 > r0)
    MFMGB_CHECK(csr_apply(ctx, A, x, epi, args, r0, r1));
  if (q1 > q0)
    MFMGB_CHECK(csr_apply(ctx, A, x, epi, args, q0, q1));
  return MFMGB_OK;
}

namespace
{
int finish_upload(mfmgb_ctx *ctx, mfmgb_csr *A)
{
  A->lanes = A->lanes_override ? A->lanes_override : choose_lanes(A->n_rows, A->nnz);
  A->device = ctx->device;
  MFMGB_CHECK(csr_measure_tiles(ctx, A));
  csr_plan_tile(A);
  return MFMGB_OK;
}

template <typename OffT>
int upload_impl(mfmgb_ctx *ctx, int64_t n_rows, int64_t n_cols, const OffT *rowptr,
                const int32_t *col, const double *val, mfmgb_csr **out)
{
  MFMGB_REQUIRE(ctx, ctx && out && rowptr && n_rows >= 0 && n_cols >= 0, "csr_upload: bad arguments");
  *out = nullptr;
  const int64_t nnz = (int64_t)rowptr[n_rows];
  MFMGB_REQUIRE(ctx, rowptr[0] == 0 && nnz >= 0, "csr_upload: rowptr must start at 0");
  MFMGB_REQUIRE(ctx, nnz == 0 || (col && val), "csr_upload: col/val are NULL");
  for (int64_t i = 0; i < n_rows; ++i)
    if (rowptr[i + 1] < rowptr[i])
      return fail(ctx, MFMGB_ERR_INVALID, "csr_upload: rowptr not monotone at row %lld", (long long)i);
  for (int64_t k = 0; k < nnz; ++k)
    if (col[k] < 0 || col[k] >= n_cols)
      return fail(ctx, MFMGB_ERR_INVALID, "csr_upload: column index %d out of range at entry %lld",
                  col[k], (long long)k);
  mfmgb_csr *A = new mfmgb_csr();
  A->n_rows = n_rows;
  A->n_cols = n_cols;
  A->nnz = nnz;
  // 64-bit row offsets from 2^31 non-zeros on (MFMGB_FORCE_OFF64=1 forces them: lets small tests cover that path)
  const char *force64 = getenv("MFMGB_FORCE_OFF64");
  A->off64 = nnz >= ((int64_t)1 << 31) - 64 || (force64 && force64[0] == '1');
  const size_t pad = 8;
  MFMGB_CUDA(ctx, cudaMalloc(&A->val, sizeof(double) * (size_t)(nnz + pad)));
  MFMGB_CUDA(ctx, cudaMalloc(&A->col, sizeof(int32_t) * (size_t)(nnz + pad)));
  MFMGB_CUDA(ctx, cudaMemsetAsync(A->val + nnz, 0, sizeof(double) * pad, ctx->stream));
  MFMGB_CUDA(ctx, cudaMemsetAsync(A->col + nnz, 0, sizeof(int32_t) * pad, ctx->stream));
  MFMGB_CUDA(ctx, cudaMemcpyAsync(A->val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
  MFMGB_CUDA(ctx, cudaMemcpyAsync(A->col, col, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
  // row offsets carry kRowptrPad entries of slack (the tile kernel copies whole 16-byte-granular chunks of them)
  const size_t rp_len = (size_t)n_rows + 1 + kRowptrPad;
  if (A->off64)
  {
    std::vector<int64_t> rp(rp_len, nnz);
    std::copy(rowptr, rowptr + n_rows + 1, rp.begin());
    MFMGB_CUDA(ctx, cudaMalloc(&A->rowptr, sizeof(int64_t) * rp_len));
    MFMGB_CUDA(ctx, cudaMemcpy(A->rowptr, rp.data(), sizeof(int64_t) * rp_len, cudaMemcpyHostToDevice));
  }
  else
  {
    std::vector<int32_t> rp(rp_len, (int32_t)nnz);
    for (int64_t i = 0; i <= n_rows; ++i)
      rp[i] = (int32_t)rowptr[i];
    MFMGB_CUDA(ctx, cudaMalloc(&A->rowptr, sizeof(int32_t) * rp_len));
    MFMGB_CUDA(ctx, cudaMemcpy(A->rowptr, rp.data(), sizeof(int32_t) * rp_len, cudaMemcpyHostToDevice));
  }
  MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  A->padded = true;
  MFMGB_CHECK(finish_upload(ctx, A));
  *out = A;
  return MFMGB_OK;
}
} // namespace
} // namespace mfmgb

extern "C"
{
  MFMGB_API int mfmgb_csr_upload(mfmgb_ctx *ctx, int64_t n_rows, int64_t n_cols, const int64_t *rowptr,
                                 const int32_t *col, const double *val, mfmgb_csr **out)
  {
    NvtxRange nvtx_range("mfmgb: operator upload");
    return upload_impl<int64_t>(ctx, n_rows, n_cols, rowptr, col, val, out);
  }

  MFMGB_API int mfmgb_csr_upload_i32(mfmgb_ctx *ctx, int64_t n_rows, int64_t n_cols, const int32_t *rowptr,
                                     const int32_t *col, const double *val, mfmgb_csr **out)
  {
    return upload_impl<int32_t>(ctx, n_rows, n_cols, rowptr, col, val, out);
  }

  MFMGB_API int mfmgb_csr_adopt_device(mfmgb_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                                       double *val_dev, int32_t *col_dev, int32_t *rowptr_dev, mfmgb_csr **out)
  {
    MFMGB_REQUIRE(ctx, ctx && out && rowptr_dev && n_rows >= 0 && n_cols >= 0 && nnz >= 0,
                  "csr_adopt_device: bad arguments");
    MFMGB_REQUIRE(ctx, nnz == 0 || (val_dev && col_dev), "csr_adopt_device: NULL arrays");
    mfmgb_csr *A = new mfmgb_csr();
    A->n_rows = n_rows;
    A->n_cols = n_cols;
    A->nnz = nnz;
    A->val = val_dev;
    A->col = col_dev;
    A->rowptr = rowptr_dev;
    A->off64 = false;
    A->owns = true;
    A->padded = false; // the caller's allocations have no slack: the tile kernel serves all but the last tile(s)
    MFMGB_CHECK(finish_upload(ctx, A));
    *out = A;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_csr_destroy(mfmgb_ctx *ctx, mfmgb_csr *A)
  {
    if (!A)
      return MFMGB_OK;
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (A->owns)
    {
      cudaFree(A->val);
      cudaFree(A->col);
      cudaFree(A->rowptr);
    }
    delete A;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_csr_info(const mfmgb_csr *A, int64_t *n_rows, int64_t *n_cols, int64_t *nnz)
  {
    if (!A)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_csr_info: A is NULL");
    if (n_rows)
      *n_rows = A->n_rows;
    if (n_cols)
      *n_cols = A->n_cols;
    if (nnz)
      *nnz = A->nnz;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_csr_device_arrays(const mfmgb_csr *A, double **val_dev, int32_t **col_dev, void **rowptr_dev,
                                        int *rowptr_is_64)
  {
    if (!A)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_csr_device_arrays: A is NULL");
    if (val_dev)
      *val_dev = A->val;
    if (col_dev)
      *col_dev = A->col;
    if (rowptr_dev)
      *rowptr_dev = A->rowptr;
    if (rowptr_is_64)
      *rowptr_is_64 = A->off64 ? 1 : 0;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_csr_download(mfmgb_ctx *ctx, const mfmgb_csr *A, int64_t *rowptr, int32_t *col, double *val)
  {
    MFMGB_REQUIRE(ctx, ctx && A && rowptr, "csr_download: bad arguments");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (A->off64)
      MFMGB_CUDA(ctx, cudaMemcpy(rowptr, A->rowptr, sizeof(int64_t) * (size_t)(A->n_rows + 1), cudaMemcpyDeviceToHost));
    else
    {
      std::vector<int32_t> rp((size_t)A->n_rows + 1);
      MFMGB_CUDA(ctx, cudaMemcpy(rp.data(), A->rowptr, sizeof(int32_t) * (size_t)(A->n_rows + 1), cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i <= A->n_rows; ++i)
        rowptr[i] = rp[i];
    }
    if (col && A->nnz)
      MFMGB_CUDA(ctx, cudaMemcpy(col, A->col, sizeof(int32_t) * (size_t)A->nnz, cudaMemcpyDeviceToHost));
    if (val && A->nnz)
      MFMGB_CUDA(ctx, cudaMemcpy(val, A->val, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost));
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_csr_transpose(mfmgb_ctx *ctx, const mfmgb_csr *A, mfmgb_csr **out)
  {
    // Setup-time operation (done once per restrictor); the reference does it on the host too
    // (EpetraExt::RowMatrix_Transpose, source/cuda/cuda_matrix_operator.cu:102-112).  Stable
    // counting sort => ascending columns, so P x_c sums in ascending coarse index: the same
    // order as the implicit Tvmult scatter of the host path.
    MFMGB_REQUIRE(ctx, ctx && A && out, "csr_transpose: bad arguments");
    std::vector<int64_t> rp((size_t)A->n_rows + 1);
    std::vector<int32_t> col((size_t)std::max<int64_t>(A->nnz, 1));
    std::vector<double> val((size_t)std::max<int64_t>(A->nnz, 1));
    MFMGB_CHECK(mfmgb_csr_download(ctx, A, rp.data(), col.data(), val.data()));
    std::vector<int64_t> trp((size_t)A->n_cols + 1, 0);
    for (int64_t k = 0; k < A->nnz; ++k)
      trp[col[k] + 1]++;
    for (int64_t c = 0; c < A->n_cols; ++c)
      trp[c + 1] += trp[c];
    std::vector<int64_t> next(trp.begin(), trp.end() - 1);
    std::vector<int32_t> tcol((size_t)std::max<int64_t>(A->nnz, 1));
    std::vector<double> tval((size_t)std::max<int64_t>(A->nnz, 1));
    for (int64_t i = 0; i < A->n_rows; ++i)
      for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
      {
        const int64_t p = next[col[k]]++;
        tcol[p] = (int32_t)i;
        tval[p] = val[k];
      }
    return mfmgb_csr_upload(ctx, A->n_cols, A->n_rows, trp.data(), tcol.data(), tval.data(), out);
  }

  MFMGB_API int mfmgb_csr_set_lanes_per_row(mfmgb_csr *A, int lanes)
  {
    if (!A || !(lanes == 0 || lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32 ||
                lanes == 256))
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_csr_set_lanes_per_row: lanes must be 0,1,2,4,8,16,32 or 256 "
                                                "(one CTA per row)");
    A->lanes_override = lanes;
    A->lanes = lanes ? lanes : choose_lanes(A->n_rows, A->nnz);
    csr_plan_tile(A);
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_csr_set_kernel(mfmgb_csr *A, int kernel)
  {
    if (!A || kernel < -1 || kernel > 1)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_csr_set_kernel: kernel must be -1 (auto), 0 (vector) or 1 (tile)");
    if (kernel == 1 && !A->tile_ok)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_csr_set_kernel: the tile-streamed kernel cannot serve this matrix "
                                                "(adopted arrays without slack, or rows too long for shared memory)");
    A->kernel_override = kernel;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_csr_get_kernel(const mfmgb_csr *A) { return A && csr_uses_tile_kernel(A) ? 1 : 0; }

  MFMGB_API int mfmgb_csr_get_lanes_per_row(const mfmgb_csr *A) { return A ? A->lanes : 0; }

  MFMGB_API int mfmgb_spmv(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, double *y)
  {
    MFMGB_REQUIRE(ctx, ctx && A && (A->n_cols == 0 || x) && (A->n_rows == 0 || y), "mfmgb_spmv: bad arguments");
    MFMGB_REQUIRE(ctx, x != y, "mfmgb_spmv: x and y must not alias");
    EpiArgs e;
    e.y = y;
    return csr_apply(ctx, A, x, Epi::Spmv, e);
  }

  MFMGB_API int mfmgb_residual_neg(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const double *b, double *r)
  {
    MFMGB_REQUIRE(ctx, ctx && A && x && b && r, "mfmgb_residual_neg: bad arguments");
    MFMGB_REQUIRE(ctx, x != r, "mfmgb_residual_neg: x and r must not alias");
    EpiArgs e;
    e.y = r;
    e.b = b;
    return csr_apply(ctx, A, x, Epi::Resid, e);
  }

  MFMGB_API int mfmgb_restrict(mfmgb_ctx *ctx, const mfmgb_csr *R, const double *r, double *b_c)
  {
    return mfmgb_spmv(ctx, R, r, b_c);
  }

  MFMGB_API int mfmgb_prolong_correct(mfmgb_ctx *ctx, const mfmgb_csr *P, const double *x_c, double *x)
  {
    MFMGB_REQUIRE(ctx, ctx && P && x_c && x, "mfmgb_prolong_correct: bad arguments");
    MFMGB_REQUIRE(ctx, x != x_c, "mfmgb_prolong_correct: x and x_c must not alias");
    EpiArgs e;
    e.y = x;
    return csr_apply(ctx, P, x_c, Epi::Sub, e);
  }
}

// dense.cu -- coarsest-level dense FP64 direct solver: factor and invert ONCE, one GEMV per solve.
//
// Replaces mfmg::CudaSolver "lu_dense" (source/cuda/cuda_solver.cu:496-515 -> lu_factorization,
// source/cuda/dealii_operator_device_helpers.cu:169-228), which runs csr2dense + getrf + getrs and
// four cudaMalloc/cudaFree on EVERY V-cycle.  Here:
//   setup  (once):  CSR -> dense, blocked right-looking LU with partial pivoting (P A = L U), the triangular
//                   factors are inverted explicitly (recursive block doubling, all GEMM), multiplied
//                   (U^-1 L^-1, only the k >= max(i,j) part of the product) and the row permutation is folded
//                   into the columns:  M = U^-1 L^-1 P  (= A^-1), one n x n row-major array.
//   solve  (hot):   x = M b  -- ONE bandwidth-bound GEMV that reads 8 n^2 bytes (the same bytes a pair of
//                   triangular applies reads) with perfectly balanced rows and no sequential dependency chain
//                   (a substitution solve needs n/nb grid-wide steps per triangle; two triangular GEMVs need
//                   three launches and have triangular load imbalance: measured 0.044 ms vs 0.022 ms at n = 4096).
//                   Multi-GPU: every rank holds M, computes a contiguous block of rows, one all-gather.
// No cuSOLVER / cuBLAS.  Tensor cores are not used (FP64, and the hot part is a GEMV).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "comm.cuh"
#include "csr.cuh"
#include "dense.cuh"

using namespace mfmgb;

namespace
{
constexpr int NB = 32; // panel width == warp size

// ---------------------------------------------------------------------------------------------
// CSR -> dense row-major (cusparseDcsr2dense at dealii_operator_device_helpers.cu:182)
// ---------------------------------------------------------------------------------------------
template <typename OffT>
__global__ void __launch_bounds__(256) csr_to_dense_kernel(int64_t n, const OffT *__restrict__ rowptr,
                                                           const int *__restrict__ col,
                                                           const double *__restrict__ val, double *__restrict__ a,
                                                           int64_t lda)
{
  // one warp per row; duplicates are summed in storage order by lane 0 only when they collide,
  // so use a serial loop per row chunk: entries of one row are distinct in practice, but keep
  // it exact for duplicates by letting a single thread own each row.
  const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (row >= n)
    return;
  for (OffT k = rowptr[row]; k < rowptr[row + 1]; ++k)
    a[row * lda + col[k]] += val[k];
}

// ---------------------------------------------------------------------------------------------
// panel factorisation: columns [k0, k0+kb), rows [k0, n); one CTA of 1024 threads (32 warps),
// lane = panel column.  getf2 with partial pivoting (first entry of maximal magnitude).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) lu_panel_kernel(double *__restrict__ a, int64_t lda, int64_t n, int64_t k0,
                                                        int kb, int *__restrict__ piv, int *__restrict__ info)
{
  __shared__ double s_val[32];
  __shared__ long long s_idx[32];
  __shared__ long long s_p;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int j = 0; j < kb; ++j)
  {
    const int64_t cj = k0 + j;
    // 1. pivot search in column cj over rows [cj, n)
    double best = -1.;
    long long bidx = 0x7fffffffffffffffLL;
    for (int64_t i = cj + tid; i < n; i += 1024)
    {
      const double v = fabs(a[i * lda + cj]);
      if (v > best)
      {
        best = v;
        bidx = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx))
      {
        best = ov;
        bidx = oi;
      }
    }
    if (lane == 0)
    {
      s_val[w] = best;
      s_idx[w] = bidx;
    }
    __syncthreads();
    if (w == 0)
    {
      best = s_val[lane];
      bidx = s_idx[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > best || (ov == best && oi < bidx))
        {
          best = ov;
          bidx = oi;
        }
      }
      if (lane == 0)
      {
        s_p = bidx;
        piv[cj] = (int)bidx;
        if (best == 0. && atomicCAS(info, 0, (int)(cj + 1)) == 0)
        {
        }
      }
      // 2. swap rows cj <-> p inside the panel
      const int64_t p = __shfl_sync(0xffffffffu, bidx, 0);
      if (p != cj && lane < kb)
      {
        const double t = a[cj * lda + k0 + lane];
        a[cj * lda + k0 + lane] = a[p * lda + k0 + lane];
        a[p * lda + k0 + lane] = t;
      }
    }
    __syncthreads();
    // 3. eliminate below the pivot inside the panel
    const double prow = lane < kb ? a[cj * lda + k0 + lane] : 0.;
    const double pivot = __shfl_sync(0xffffffffu, prow, j);
    if (pivot != 0.)
    {
      const double inv = 1. / pivot;
      for (int64_t i = cj + 1 + w; i < n; i += 32)
      {
        double v = lane < kb ? a[i * lda + k0 + lane] : 0.;
        const double l = __dmul_rn(__shfl_sync(0xffffffffu, v, j), inv);
        if (lane == j)
          v = l;
        else if (lane > j)
          v = fma(-l, prow, v);
        if (lane >= j && lane < kb)
          a[i * lda + k0 + lane] = v;
      }
    }
    __syncthreads();
  }
}

// apply the panel's row interchanges to the columns outside the panel
__global__ void __launch_bounds__(256) lu_swap_kernel(double *__restrict__ a, int64_t lda, int64_t n, int64_t k0, int kb,
                                                      const int *__restrict__ piv)
{
  int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= n - kb)
    return;
  if (c >= k0)
    c += kb; // skip the panel's own columns
  for (int j = 0; j < kb; ++j)
  {
    const int64_t r = k0 + j, p = piv[r];
    if (p != r)
    {
      const double t = a[r * lda + c];
      a[r * lda + c] = a[p * lda + c];
      a[p * lda + c] = t;
    }
  }
}

// U12 = L11^-1 A12 : one thread per column right of the panel
__global__ void __launch_bounds__(256) lu_trsm_kernel(double *__restrict__ a, int64_t lda, int64_t n, int64_t k0, int kb)
{
  __shared__ double L[NB][NB + 1];
  for (int e = threadIdx.x; e < kb * kb; e += 256)
    L[e / kb][e % kb] = a[(k0 + e / kb) * lda + k0 + e % kb];
  __syncthreads();
  const int64_t c = k0 + kb + (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= n)
    return;
  double u[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i)
    u[i] = i < kb ? a[(k0 + i) * lda + c] : 0.;
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int i = j + 1; i < NB; ++i)
      if (i < kb)
        u[i] = fma(-L[i][j], u[j], u[i]);
#pragma unroll
  for (int i = 0; i < NB; ++i)
    if (i < kb)
      a[(k0 + i) * lda + c] = u[i];
}

// ---------------------------------------------------------------------------------------------
// generic FP64 GEMM tile: C[64x64 tile] = alpha * A[M x K] B[K x N] + beta * C, row-major.
// 256 threads, 4x4 outputs per thread, K in chunks of 16 through shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ void gemm_tile(int64_t M, int64_t N, int64_t K, const double *__restrict__ A, int64_t lda,
                                          const double *__restrict__ B, int64_t ldb, double *__restrict__ C,
                                          int64_t ldc, double alpha, double beta, int64_t tile_m, int64_t tile_n,
                                          int64_t k_begin = 0)
{
  __shared__ double As[TK][TM + 1];
  __shared__ double Bs[TK][TN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = tile_m * TM, n0 = tile_n * TN;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc[i][j] = 0.;
  for (int64_t k0 = k_begin; k0 < K; k0 += TK)
  {
    // A tile: 64 rows x 16 k  (each thread loads 4 elements)
#pragma unroll
    for (int e = 0; e < 4; ++e)
    {
      const int idx = tid + e * 256; // 0..1023
      const int r = idx >> 4, kk = idx & 15;
      const int64_t gr = m0 + r, gk = k0 + kk;
      As[kk][r] = (gr < M && gk < K) ? A[gr * lda + gk] : 0.;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
    {
      const int idx = tid + e * 256;
      const int kk = idx >> 6, c = idx & 63;
      const int64_t gk = k0 + kk, gc = n0 + c;
      Bs[kk][c] = (gk < K && gc < N) ? B[gk * ldb + gc] : 0.;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk)
    {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    const int64_t gr = m0 + ty * 4 + i;
    if (gr >= M)
      continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const int64_t gc = n0 + tx + 16 * j;
      if (gc < N)
      {
        double *p = C + gr * ldc + gc;
        *p = beta == 0. ? alpha * acc[i][j] : fma(alpha, acc[i][j], beta * *p);
      }
    }
  }
}

// C = alpha A B + beta C (row-major, any sizes)
__global__ void __launch_bounds__(256) gemm_kernel(int64_t M, int64_t N, int64_t K, const double *__restrict__ A,
                                                   int64_t lda, const double *__restrict__ B, int64_t ldb,
                                                   double *__restrict__ C, int64_t ldc, double alpha, double beta)
{
  gemm_tile(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, blockIdx.y, blockIdx.x);
}

// trailing update A22 -= L21 U12
__global__ void __launch_bounds__(256) lu_update_kernel(double *__restrict__ a, int64_t lda, int64_t n, int64_t k0, int kb)
{
  const int64_t k1 = k0 + kb;
  const int64_t M = n - k1;
  gemm_tile(M, M, kb, a + k1 * lda + k0, lda, a + k0 * lda + k1, lda, a + k1 * lda + k1, lda, -1., 1., blockIdx.y,
            blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// triangular inverses
// ---------------------------------------------------------------------------------------------
// base case: invert the 32x32 diagonal blocks.  One warp per block, lane = column of the inverse.
__global__ void __launch_bounds__(32) tri_diag_inverse_kernel(const double *__restrict__ lu, int64_t lda, int64_t n,
                                                              double *__restrict__ linv, double *__restrict__ uinv)
{
  __shared__ double S[NB][NB + 1];
  const int64_t r0 = (int64_t)blockIdx.x * NB;
  const int m = (int)min((int64_t)NB, n - r0);
  const int j = threadIdx.x;
  for (int i = 0; i < m; ++i)
    if (j < m)
      S[i][j] = lu[(r0 + i) * lda + r0 + j];
  __syncwarp();
  if (j < m)
  {
    double x[NB];
    // unit lower: X = L^-1, column j
#pragma unroll
    for (int i = 0; i < NB; ++i)
      x[i] = 0.;
#pragma unroll
    for (int i = 0; i < NB; ++i)
    {
      if (i == j)
        x[i] = 1.;
      else if (i > j && i < m)
      {
        double s = 0.;
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k >= j && k < i)
            s = fma(S[i][k], x[k], s);
        x[i] = -s;
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (i < m)
        linv[(r0 + i) * lda + r0 + j] = i >= j ? x[i] : 0.;
    // upper: X = U^-1, column j
#pragma unroll
    for (int i = 0; i < NB; ++i)
      x[i] = 0.;
#pragma unroll
    for (int ii = 0; ii < NB; ++ii)
    {
      const int i = NB - 1 - ii;
      if (i == j)
        x[i] = 1. / S[i][i];
      else if (i < j)
      {
        double s = 0.;
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k > i && k <= j)
            s = fma(S[i][k], x[k], s);
        x[i] = -s / S[i][i];
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (i < m)
        uinv[(r0 + i) * lda + r0 + j] = i <= j ? x[i] : 0.;
  }
}

// level step for block size s: pair p covers [r0, mid) and [mid, r1).
//   LOWER: T = L21 X11        then X21 = -X22 T        (X = L^-1)
//   UPPER: T = U12 X22        then X12 = -X11 T        (X = U^-1)
template <bool LOWER, int STEP>
__global__ void __launch_bounds__(256) tri_level_kernel(const double *__restrict__ lu, double *__restrict__ X,
                                                        double *__restrict__ T, int64_t lda, int64_t n, int64_t s)
{
  const int64_t p = blockIdx.z;
  const int64_t r0 = p * 2 * s;
  const int64_t mid = min(r0 + s, n), r1 = min(r0 + 2 * s, n);
  const int64_t m1 = mid - r0, m2 = r1 - mid;
  if (m2 <= 0)
    return;
  if (LOWER)
  {
    // off-diagonal block lives at rows [mid,r1), cols [r0,mid): m2 x m1
    if ((int64_t)blockIdx.y * TM >= m2 || (int64_t)blockIdx.x * TN >= m1)
      return;
    if (STEP == 1)
      gemm_tile(m2, m1, m1, lu + mid * lda + r0, lda, X + r0 * lda + r0, lda, T + mid * lda + r0, lda, 1., 0.,
                blockIdx.y, blockIdx.x);
    else
      gemm_tile(m2, m1, m2, X + mid * lda + mid, lda, T + mid * lda + r0, lda, X + mid * lda + r0, lda, -1., 0.,
                blockIdx.y, blockIdx.x);
  }
  else
  {
    // off-diagonal block lives at rows [r0,mid), cols [mid,r1): m1 x m2
    if ((int64_t)blockIdx.y * TM >= m1 || (int64_t)blockIdx.x * TN >= m2)
      return;
    if (STEP == 1)
      gemm_tile(m1, m2, m2, lu + r0 * lda + mid, lda, X + mid * lda + mid, lda, T + r0 * lda + mid, lda, 1., 0.,
                blockIdx.y, blockIdx.x);
    else
      gemm_tile(m1, m2, m1, X + r0 * lda + r0, lda, T + r0 * lda + mid, lda, X + r0 * lda + mid, lda, -1., 0.,
                blockIdx.y, blockIdx.x);
  }
}

// Inv = U^-1 L^-1: entry (i, j) only sums over k >= max(i, j) (U^-1 is upper, L^-1 lower triangular)
__global__ void __launch_bounds__(256) inv_product_kernel(const double *__restrict__ uinv,
                                                          const double *__restrict__ linv, double *__restrict__ out,
                                                          int64_t lda, int64_t n)
{
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
  const int64_t kb = (m0 > n0 ? m0 : n0) / TK * TK;
  gemm_tile(n, n, n, uinv, lda, linv, lda, out, lda, 1., 0., blockIdx.y, blockIdx.x, kb);
}

// M[i][perm[k]] = Inv[i][k]  (x = Inv (P b) with (P b)[k] = b[perm[k]]); padding columns are zeroed
__global__ void __launch_bounds__(256) fold_perm_kernel(const double *__restrict__ inv, const int *__restrict__ perm,
                                                        double *__restrict__ out, int64_t lda, int64_t n)
{
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (k < n)
    out[i * lda + perm[k]] = inv[i * lda + k];
  else if (k < lda)
    out[i * lda + k] = 0.;
}

// ---------------------------------------------------------------------------------------------
// solve: x = M b.  TPR threads per row, 128-bit loads, four independent accumulators per thread and a fixed
// reduction tree (deterministic).  Rows [row0, row0 + n_out) go to out[0 .. n_out); rows >= n give 0.
// ---------------------------------------------------------------------------------------------
template <int TPR>
__global__ void __launch_bounds__(256) gemv_kernel(int64_t n, const double *__restrict__ M, int64_t lda,
                                                   const double *__restrict__ v, double *__restrict__ out,
                                                   int64_t row0, int64_t n_out)
{
  __shared__ double sm[8];
  constexpr int ROWS_PER_BLOCK = 256 / TPR;
  const int t = threadIdx.x % TPR;
  const int64_t slot = (int64_t)blockIdx.x * ROWS_PER_BLOCK + threadIdx.x / TPR;
  const int64_t i = row0 + slot;
  double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
  if (slot < n_out && i < n)
  {
    // lda is a multiple of 4 and the padding columns of M are zero: whole rows in 128-bit pieces
    const double2 *row2 = reinterpret_cast<const double2 *>(M + i * lda);
    const double2 *v2 = reinterpret_cast<const double2 *>(v);
    const int64_t npairs = n >> 1;
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
    if ((n & 1) && t == 0)
      s1 = fma(M[i * lda + n - 1], v[n - 1], s1);
  }
  double s = (s0 + s1) + (s2 + s3);
  if (TPR == 256)
  {
    s = block_sum<256>(s, sm);
    if (threadIdx.x == 0 && slot < n_out)
      out[slot] = s;
  }
  else
  {
    s = subwarp_sum<(TPR < 32 ? TPR : 32)>(s);
    if (t == 0 && slot < n_out)
      out[slot] = s;
  }
}

// diagonal of the packed factors (pivots of U)
__global__ void __launch_bounds__(256) lu_diag_kernel(const double *__restrict__ lu, int64_t lda, int64_t n,
                                                      double *__restrict__ diag)
{
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n)
    diag[i] = lu[i * lda + i];
}

// x = U^-1 L^-1 P b by substitution, in the operation order of getrs' reference implementation (the oracle's
// orc_lu_solve: row-oriented, ascending column index, separate multiply and subtract): one thread, small systems
__global__ void lu_subst_serial_kernel(const double *__restrict__ lu, int64_t lda, int64_t n, const int *__restrict__ perm,
                                       const double *__restrict__ b, double *__restrict__ x)
{
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  for (int64_t i = 0; i < n; ++i)
  {
    double s = b[perm[i]];
    const double *row = lu + i * lda;
    for (int64_t j = 0; j < i; ++j)
      s = __dsub_rn(s, __dmul_rn(row[j], x[j]));
    x[i] = s;
  }
  for (int64_t i = n - 1; i >= 0; --i)
  {
    double s = x[i];
    const double *row = lu + i * lda;
    for (int64_t j = i + 1; j < n; ++j)
      s = __dsub_rn(s, __dmul_rn(row[j], x[j]));
    x[i] = __ddiv_rn(s, row[i]);
  }
}

// the same solve, column-oriented, one CTA (larger ill-conditioned systems): every entry still receives its updates in
// ascending k in the forward sweep; the backward sweep runs in descending k
__global__ void __launch_bounds__(1024)
    lu_subst_cta_kernel(const double *__restrict__ lu, int64_t lda, int64_t n, const int *__restrict__ perm,
                        const double *__restrict__ b, double *__restrict__ x)
{
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
    x[i] = b[perm[i]];
  __syncthreads();
  for (int64_t k = 0; k < n; ++k)
  {
    const double xk = x[k];
    for (int64_t i = k + 1 + threadIdx.x; i < n; i += blockDim.x)
      x[i] = __dsub_rn(x[i], __dmul_rn(lu[i * lda + k], xk));
    __syncthreads();
  }
  for (int64_t k = n - 1; k >= 0; --k)
  {
    if (threadIdx.x == 0)
      x[k] = __ddiv_rn(x[k], lu[k * lda + k]);
    __syncthreads();
    const double xk = x[k];
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x)
      x[i] = __dsub_rn(x[i], __dmul_rn(lu[i * lda + k], xk));
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// solve, large n: TMA-streamed GEMV.  The round-1 kernel (one CTA per row, direct 128-bit loads) reached 4.1 TB/s at
// n_c = 4096: 4096 short-lived CTAs, each exposing its own ramp-up and a block reduction.  Here the rows of M do not
// occupy warps while they travel: one persistent CTA per SM owns a contiguous run of rows and walks the matrix in
// column chunks of kGemvChunk doubles; one elected producer thread streams (row, chunk) pieces into a shared-memory
// ring (one small ring per consumer warp) with cp.async.bulk (1-D TMA, L2 evict-first) guarded by full / empty
// mbarriers; consumer warp w takes the pieces of its rows (row mod 8 == w), multiplies them with the chunk of b held in
// shared memory (loaded once per
// chunk and CTA, not once per row), reduces with a fixed shuffle tree and adds the result to the row's accumulator in
// shared memory -- only that warp ever touches it, so the only block-wide synchronisation is one consumer barrier per
// column chunk.  Summation order depends on sizes only: deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int kGemvChunk = 1024;  // doubles per piece (8 KB)
constexpr int kGemvDepth = 3;     // every consumer warp has its OWN ring of this depth (single producer, single consumer:
                                  // with one shared ring a fast warp could wait on a slot use that is two phases ahead of
                                  // the barrier, which mbarrier parity waits cannot tell from a completed one)
constexpr int kGemvStages = 8 * kGemvDepth;
constexpr int kGemvWarps = 8;
constexpr int kGemvThreads = (kGemvWarps + 1) * 32;
constexpr int kGemvRowsMax = 256; // rows per CTA (accumulators in shared memory)

__device__ __forceinline__ uint32_t gv_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gv_mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gv_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gv_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gv_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gv_mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gv_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gv_mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t done;
  do
  {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}"
                 : "=r"(done)
                 : "r"(gv_smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}

__global__ void __launch_bounds__(kGemvThreads, 1)
    gemv_stream_kernel(int64_t n, const double *__restrict__ M, int64_t lda, const double *__restrict__ v,
                       double *__restrict__ out, int64_t row0, int64_t n_out)
{
  extern __shared__ __align__(128) unsigned char gv_smem[];
  double *ring = reinterpret_cast<double *>(gv_smem);           // [stages][chunk]
  double *vbuf = ring + (size_t)kGemvStages * kGemvChunk;       // [2][chunk]
  double *acc = vbuf + 2 * kGemvChunk;                          // [kGemvRowsMax]
  uint64_t *full = reinterpret_cast<uint64_t *>(acc + kGemvRowsMax);
  uint64_t *empty = full + kGemvStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0)
  {
    for (int s = 0; s < kGemvStages; ++s)
    {
      gv_mbar_init(full + s, 1);
      gv_mbar_init(empty + s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // this CTA's contiguous run of output slots; its valid rows (< n) are a prefix of the run, the others give 0
  const int64_t rpc = (n_out + gridDim.x - 1) / gridDim.x;
  const int64_t s_begin = (int64_t)blockIdx.x * rpc;
  const int64_t s_end = s_begin + rpc < n_out ? s_begin + rpc : n_out;
  int64_t s_valid = s_end;
  if (row0 + s_valid > n)
    s_valid = n - row0 > s_begin ? n - row0 : s_begin;
  const int n_rows = (int)(s_valid > s_begin ? s_valid - s_begin : 0); // <= kGemvRowsMax
  const int n_chunks = (int)((lda + kGemvChunk - 1) / kGemvChunk);
  if (warp == kGemvWarps)
  {
    // ---- producer: pieces in (chunk, row) order ----
    if (lane != 0)
      return;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    int cnt[kGemvWarps];
#pragma unroll
    for (int w = 0; w < kGemvWarps; ++w)
      cnt[w] = 0;
    for (int c = 0; c < n_chunks; ++c)
    {
      const int64_t len = lda - (int64_t)c * kGemvChunk < kGemvChunk ? lda - (int64_t)c * kGemvChunk : kGemvChunk;
      for (int r0 = 0; r0 < n_rows; r0 += kGemvWarps)
      {
#pragma unroll
        for (int w = 0; w < kGemvWarps; ++w)
        {
          const int r = r0 + w;
          if (r >= n_rows)
            break;
          const int st = w * kGemvDepth + cnt[w] % kGemvDepth;
          const uint32_t ph = (uint32_t)((cnt[w] / kGemvDepth) & 1);
          ++cnt[w];
          gv_mbar_wait(empty + st, ph ^ 1u);
          gv_mbar_expect_tx(full + st, (uint32_t)(len * 8));
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
              :
              : "r"(gv_smem_u32(ring + (size_t)st * kGemvChunk)),
                "l"(M + (row0 + s_begin + r) * lda + (int64_t)c * kGemvChunk), "r"((uint32_t)(len * 8)),
                "r"(gv_smem_u32(full + st)), "l"(policy)
              : "memory");
        }
      }
    }
    return;
  }
  // ---- consumers ----
  const int ct = threadIdx.x; // 0 .. 255
  int my_cnt = 0;             // pieces this warp has taken from its ring
  for (int r = ct; r < kGemvRowsMax; r += kGemvWarps * 32)
    acc[r] = 0.;
  for (int c = 0; c < n_chunks; ++c)
  {
    const int64_t len = lda - (int64_t)c * kGemvChunk < kGemvChunk ? lda - (int64_t)c * kGemvChunk : kGemvChunk;
    // the chunk of b for these columns (zero beyond n: the padding columns of M are zero as well)
    double *vb = vbuf + (size_t)(c & 1) * kGemvChunk;
    for (int j = ct; j < kGemvChunk; j += kGemvWarps * 32)
    {
      const int64_t col = (int64_t)c * kGemvChunk + j;
      vb[j] = col < n ? v[col] : 0.;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kGemvWarps * 32) : "memory"); // consumers only
    for (int r = warp; r < n_rows; r += kGemvWarps)
    {
      const int st = warp * kGemvDepth + my_cnt % kGemvDepth; // this warp's own ring
      const uint32_t ph = (uint32_t)((my_cnt / kGemvDepth) & 1);
      ++my_cnt;
      gv_mbar_wait(full + st, ph);
      const double *m = ring + (size_t)st * kGemvChunk;
      double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
      // 128-bit accesses, lane-contiguous: element pairs lane * 2 + 64 k
#pragma unroll 4
      for (int j = lane * 2; j < (int)len; j += 128)
      {
        const double2 a0 = *reinterpret_cast<const double2 *>(m + j), b0 = *reinterpret_cast<const double2 *>(vb + j);
        s0 = fma(a0.x, b0.x, s0);
        s1 = fma(a0.y, b0.y, s1);
        if (j + 64 < (int)len)
        {
          const double2 a1 = *reinterpret_cast<const double2 *>(m + j + 64),
                        b1 = *reinterpret_cast<const double2 *>(vb + j + 64);
          s2 = fma(a1.x, b1.x, s2);
          s3 = fma(a1.y, b1.y, s3);
        }
      }
      __syncwarp();
      if (lane == 0)
        gv_mbar_arrive(empty + st); // the piece is in registers: the slot may be refilled
      const double w = subwarp_sum<32>((s0 + s1) + (s2 + s3));
      if (lane == 0)
        acc[r] += w; // only this warp touches acc[r]; chunks are added in ascending order
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kGemvWarps * 32) : "memory");
  for (int64_t slot = s_begin + ct; slot < s_end; slot += kGemvWarps * 32)
    out[slot] = slot < s_valid ? acc[slot - s_begin] : 0.;
}

int launch_gemv_stream(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *out, int64_t row0, int64_t n_out)
{
  const size_t smem = sizeof(double) * ((size_t)(kGemvStages + 2) * kGemvChunk + kGemvRowsMax) +
                      sizeof(uint64_t) * 2 * kGemvStages;
  static unsigned long long configured = 0; // one bit per device
  if (!((configured >> (ctx->device & 63)) & 1ull))
  {
    MFMGB_CUDA(ctx, cudaFuncSetAttribute(gemv_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    configured |= 1ull << (ctx->device & 63);
  }
  // one CTA per SM; more only when a CTA would own more rows than it has accumulators for
  const int64_t grid = std::max<int64_t>(std::min<int64_t>(n_out, ctx->num_sms), ceil_div(n_out, kGemvRowsMax));
  gemv_stream_kernel<<<(unsigned)grid, kGemvThreads, smem, ctx->stream>>>(D->n, D->inv, D->lda, b, out, row0, n_out);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

int launch_gemv(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *out, int64_t row0, int64_t n_out)
{
  // large operators: the TMA-streamed kernel (MFMGB_GEMV_STREAM=0 keeps the direct-load kernel: measurement aid)
  static const bool stream = [] {
    const char *v = getenv("MFMGB_GEMV_STREAM");
    return !(v && v[0] == '0');
  }();
  if (stream && !D->direct_gemv && D->n >= 1024 && n_out >= 256)
    return launch_gemv_stream(ctx, D, b, out, row0, n_out);

  if (n_out <= 0)
    return MFMGB_OK;
  if (D->n >= 1024)
    gemv_kernel<256><<<(unsigned)n_out, 256, 0, ctx->stream>>>(D->n, D->inv, D->lda, b, out, row0, n_out);
  else
    gemv_kernel<32><<<(unsigned)ceil_div(n_out, 8), 256, 0, ctx->stream>>>(D->n, D->inv, D->lda, b, out, row0, n_out);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}
} // namespace

namespace mfmgb
{
int dense_solve_async(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *x)
{
  const int64_t n = D->n;
  if (n == 0)
    return MFMGB_OK;
  if (b == x || (reinterpret_cast<uintptr_t>(b) & 15))
  {
    // in-place solve, or a right-hand side that is not 16-byte aligned: stage it (the GEMV reads b with 128-bit loads)
    MFMGB_CUDA(ctx, cudaMemcpyAsync(D->work0, b, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    b = D->work0;
  }
  if (D->substitution && !D->distributed)
  {
    if (n <= 128)
      lu_subst_serial_kernel<<<1, 32, 0, ctx->stream>>>(D->lu, D->lda, n, D->perm, b, x);
    else
      lu_subst_cta_kernel<<<1, 1024, 0, ctx->stream>>>(D->lu, D->lda, n, D->perm, b, x);
    MFMGB_LAUNCHED(ctx);
    return MFMGB_OK;
  }
  if (!D->distributed)
    return launch_gemv(ctx, D, b, x, 0, n);
  // every rank: its block of rows of x = M b, then one all-gather (b and M are replicated)
  mfmgb_comm *c = ctx_comm(ctx);
  MFMGB_CHECK(launch_gemv(ctx, D, b, D->chunk, (int64_t)D->rank * D->rows_per_rank, D->rows_per_rank));
  MFMGB_NCCL(ctx, ncclAllGather(D->chunk, D->gathered, (size_t)D->rows_per_rank, ncclDouble, c->nccl, ctx->stream));
  MFMGB_CUDA(ctx, cudaMemcpyAsync(x, D->gathered, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  return MFMGB_OK;
}

int dense_enable_distributed(mfmgb_ctx *ctx, mfmgb_dense *D)
{
  mfmgb_comm *c = ctx_comm(ctx);
  if (!c || c->nranks < 2 || D->n == 0)
    return MFMGB_OK;
  D->nranks = c->nranks;
  D->rank = c->rank;
  D->rows_per_rank = ceil_div(D->n, (int64_t)c->nranks);
  MFMGB_CUDA(ctx, cudaMalloc(&D->chunk, sizeof(double) * (size_t)D->rows_per_rank));
  MFMGB_CUDA(ctx, cudaMalloc(&D->gathered, sizeof(double) * (size_t)(D->rows_per_rank * c->nranks)));
  D->distributed = true;
  return MFMGB_OK;
}
} // namespace mfmgb

namespace mfmgb
{
int dense_gemm(mfmgb_ctx *ctx, int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
               int64_t ldb, double *C, int64_t ldc, double alpha, double beta)
{
  if (M <= 0 || N <= 0)
    return MFMGB_OK;
  dim3 grid((unsigned)ceil_div(N, TN), (unsigned)ceil_div(M, TM));
  gemm_kernel<<<grid, 256, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

int dense_apply_rows(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *out, int64_t row0, int64_t n_out)
{
  return launch_gemv(ctx, D, b, out, row0, n_out);
}

// dense row-major copy (n_rows x lda, zero padded) of a CSR matrix; duplicates are summed
int csr_to_dense_device(mfmgb_ctx *ctx, const mfmgb_csr *A, int64_t lda, double **out)
{
  const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(A->n_rows * lda, 1);
  double *a = nullptr;
  MFMGB_CUDA(ctx, cudaMalloc(&a, bytes));
  MFMGB_CUDA(ctx, cudaMemsetAsync(a, 0, bytes, ctx->stream));
  if (A->n_rows > 0)
  {
    const unsigned nb = (unsigned)ceil_div(A->n_rows, 256);
    if (A->off64)
      csr_to_dense_kernel<int64_t><<<nb, 256, 0, ctx->stream>>>(A->n_rows, (const int64_t *)A->rowptr, A->col, A->val, a, lda);
    else
      csr_to_dense_kernel<int32_t><<<nb, 256, 0, ctx->stream>>>(A->n_rows, (const int32_t *)A->rowptr, A->col, A->val, a, lda);
    MFMGB_LAUNCHED(ctx);
  }
  *out = a;
  return MFMGB_OK;
}

// factorise / invert the dense row-major matrix `lu` (n x lda, lda = n rounded up to 4, padding columns zero) in place;
// takes ownership of `lu` (it becomes D->inv)
int dense_factor_device(mfmgb_ctx *ctx, double *lu, int64_t n, mfmgb_dense **out)
{
    *out = nullptr;
    const int64_t lda = (n + 3) & ~(int64_t)3;
    mfmgb_dense *D = new mfmgb_dense();
    D->n = n;
    D->lda = lda;
    const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(n * lda, 1);
    double *linv = nullptr, *uinv = nullptr, *T = nullptr;
    int *piv = nullptr, *info_dev = nullptr;
    MFMGB_CUDA(ctx, cudaMalloc(&linv, bytes));
    MFMGB_CUDA(ctx, cudaMalloc(&uinv, bytes));
    MFMGB_CUDA(ctx, cudaMalloc(&T, bytes));
    MFMGB_CUDA(ctx, cudaMalloc(&piv, sizeof(int) * (size_t)std::max<int64_t>(n, 1)));
    MFMGB_CUDA(ctx, cudaMalloc(&info_dev, sizeof(int)));
    MFMGB_CUDA(ctx, cudaMalloc(&D->perm, sizeof(int) * (size_t)std::max<int64_t>(n, 1)));
    MFMGB_CUDA(ctx, cudaMalloc(&D->work0, sizeof(double) * (size_t)(n + 2)));
    MFMGB_CUDA(ctx, cudaMalloc(&D->work1, sizeof(double) * (size_t)(n + 2)));
    MFMGB_CUDA(ctx, cudaMemsetAsync(D->work0, 0, sizeof(double) * (size_t)(n + 2), ctx->stream));
    MFMGB_CUDA(ctx, cudaMemsetAsync(D->work1, 0, sizeof(double) * (size_t)(n + 2), ctx->stream));
    MFMGB_CUDA(ctx, cudaMemsetAsync(linv, 0, bytes, ctx->stream));
    MFMGB_CUDA(ctx, cudaMemsetAsync(uinv, 0, bytes, ctx->stream));
    MFMGB_CUDA(ctx, cudaMemsetAsync(T, 0, bytes, ctx->stream));
    MFMGB_CUDA(ctx, cudaMemsetAsync(info_dev, 0, sizeof(int), ctx->stream));
    cudaStream_t st = ctx->stream;
    if (n > 0)
    {
      // blocked right-looking LU
      for (int64_t k0 = 0; k0 < n; k0 += NB)
      {
        const int kb = (int)std::min<int64_t>(NB, n - k0);
        lu_panel_kernel<<<1, 1024, 0, st>>>(lu, lda, n, k0, kb, piv, info_dev);
        MFMGB_LAUNCHED(ctx);
        if (n - kb > 0)
        {
          lu_swap_kernel<<<(unsigned)ceil_div(n - kb, 256), 256, 0, st>>>(lu, lda, n, k0, kb, piv);
          MFMGB_LAUNCHED(ctx);
        }
        const int64_t rest = n - k0 - kb;
        if (rest > 0)
        {
          lu_trsm_kernel<<<(unsigned)ceil_div(rest, 256), 256, 0, st>>>(lu, lda, n, k0, kb);
          MFMGB_LAUNCHED(ctx);
          dim3 grid((unsigned)ceil_div(rest, TN), (unsigned)ceil_div(rest, TM));
          lu_update_kernel<<<grid, 256, 0, st>>>(lu, lda, n, k0, kb);
          MFMGB_LAUNCHED(ctx);
        }
      }
      int info = 0;
      MFMGB_CUDA(ctx, cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
      std::vector<int> hpiv((size_t)n);
      MFMGB_CUDA(ctx, cudaMemcpyAsync(hpiv.data(), piv, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
      MFMGB_CUDA(ctx, cudaStreamSynchronize(st));
      if (info != 0)
      {
        cudaFree(lu);
        cudaFree(linv);
        cudaFree(uinv);
        cudaFree(T);
        cudaFree(piv);
        cudaFree(info_dev);
        cudaFree(D->perm);
        cudaFree(D->work0);
        cudaFree(D->work1);
        delete D;
        return fail(ctx, MFMGB_ERR_SINGULAR, "mfmgb_dense_factor: zero pivot at column %d", info - 1);
      }
      std::vector<int> perm((size_t)n);
      for (int64_t i = 0; i < n; ++i)
        perm[i] = (int)i;
      int64_t swaps = 0;
      for (int64_t k = 0; k < n; ++k)
        if (hpiv[k] != k)
        {
          std::swap(perm[k], perm[hpiv[k]]);
          ++swaps;
        }
      D->num_swaps = swaps;
      MFMGB_CUDA(ctx, cudaMemcpyAsync(D->perm, perm.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
      MFMGB_CUDA(ctx, cudaStreamSynchronize(st));
      // conditioning from the pivots: a spread of more than 12 decades keeps the factors and solves by substitution
      {
        std::vector<double> hdiag((size_t)n);
        lu_diag_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(lu, lda, n, D->work0);
        MFMGB_LAUNCHED(ctx);
        MFMGB_CUDA(ctx, cudaMemcpyAsync(hdiag.data(), D->work0, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
        MFMGB_CUDA(ctx, cudaStreamSynchronize(st));
        double dmin = std::fabs(hdiag[0]), dmax = dmin;
        for (double v : hdiag)
        {
          dmin = std::min(dmin, std::fabs(v));
          dmax = std::max(dmax, std::fabs(v));
        }
        D->pivot_ratio = dmax > 0. ? dmin / dmax : 0.;
        const char *force = getenv("MFMGB_DENSE_SOLVE"); // "substitution" / "inverse" override the automatic choice
        D->substitution = force ? !strcmp(force, "substitution") : D->pivot_ratio < 1e-12;
        if (D->substitution)
        {
          MFMGB_CUDA(ctx, cudaMalloc(&D->lu, bytes));
          MFMGB_CUDA(ctx, cudaMemcpyAsync(D->lu, lu, bytes, cudaMemcpyDeviceToDevice, st));
        }
      }
      // triangular inverses by block doubling
      const int64_t nblk = ceil_div(n, NB);
      tri_diag_inverse_kernel<<<(unsigned)nblk, 32, 0, st>>>(lu, lda, n, linv, uinv);
      MFMGB_LAUNCHED(ctx);
      for (int64_t s = NB; s < n; s *= 2)
      {
        const int64_t npairs = ceil_div(n, 2 * s);
        dim3 grid((unsigned)ceil_div(s, TN), (unsigned)ceil_div(s, TM), (unsigned)npairs);
        tri_level_kernel<true, 1><<<grid, 256, 0, st>>>(lu, linv, T, lda, n, s);
        MFMGB_LAUNCHED(ctx);
        tri_level_kernel<true, 2><<<grid, 256, 0, st>>>(lu, linv, T, lda, n, s);
        MFMGB_LAUNCHED(ctx);
        tri_level_kernel<false, 1><<<grid, 256, 0, st>>>(lu, uinv, T, lda, n, s);
        MFMGB_LAUNCHED(ctx);
        tri_level_kernel<false, 2><<<grid, 256, 0, st>>>(lu, uinv, T, lda, n, s);
        MFMGB_LAUNCHED(ctx);
      }
      // M = U^-1 L^-1 P into `lu` (the factors themselves are not needed any more)
      dim3 ggrid((unsigned)ceil_div(n, TN), (unsigned)ceil_div(n, TM));
      inv_product_kernel<<<ggrid, 256, 0, st>>>(uinv, linv, T, lda, n);
      MFMGB_LAUNCHED(ctx);
      dim3 pgrid((unsigned)ceil_div(lda, 256), (unsigned)n);
      fold_perm_kernel<<<pgrid, 256, 0, st>>>(T, D->perm, lu, lda, n);
      MFMGB_LAUNCHED(ctx);
      MFMGB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    D->inv = lu;
    cudaFree(linv);
    cudaFree(uinv);
    cudaFree(T);
    cudaFree(piv);
    cudaFree(info_dev);
    *out = D;
    return MFMGB_OK;
  }

} // namespace mfmgb

extern "C"
{
  MFMGB_API int mfmgb_dense_factor(mfmgb_ctx *ctx, const mfmgb_csr *A, mfmgb_dense **out)
  {
    NvtxRange nvtx_range("mfmgb: coarse solver setup (dense LU)");
    MFMGB_REQUIRE(ctx, ctx && A && out, "mfmgb_dense_factor: bad arguments");
    MFMGB_REQUIRE(ctx, A->n_rows == A->n_cols, "mfmgb_dense_factor: the matrix is not square");
    *out = nullptr;
    const int64_t n = A->n_rows;
    const int64_t lda = (n + 3) & ~(int64_t)3;
    double *lu = nullptr;
    MFMGB_CHECK(csr_to_dense_device(ctx, A, lda, &lu));
    return dense_factor_device(ctx, lu, n, out);
  }

  MFMGB_API int mfmgb_dense_destroy(mfmgb_ctx *ctx, mfmgb_dense *D)
  {
    if (!D)
      return MFMGB_OK;
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(D->inv);
    cudaFree(D->lu);
    cudaFree(D->perm);
    cudaFree(D->work0);
    cudaFree(D->work1);
    cudaFree(D->chunk);
    cudaFree(D->gathered);
    delete D;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_dense_solve(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *x)
  {
    MFMGB_REQUIRE(ctx, ctx && D && (D->n == 0 || (b && x)), "mfmgb_dense_solve: bad arguments");
    return dense_solve_async(ctx, D, b, x);
  }

  MFMGB_API int64_t mfmgb_dense_size(const mfmgb_dense *D) { return D ? D->n : 0; }
  MFMGB_API int64_t mfmgb_dense_num_swaps(const mfmgb_dense *D) { return D ? D->num_swaps : 0; }
  MFMGB_API int mfmgb_dense_solve_mode(const mfmgb_dense *D, double *pivot_ratio)
  {
    if (D && pivot_ratio)
      *pivot_ratio = D->pivot_ratio;
    return D && D->substitution ? 1 : 0;
  }
}
